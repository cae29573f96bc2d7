"""Drop-in for the reference's k-means-color-clustering/drawGridsAndOutputCSVChange.py: the grid/CSV
stage with the flow computed inline and every cell ROI written to
``OutImgs/<video>/<frame>/<cell>.png`` (:107-109) -- the script that produced the reference's PNG
fixtures.

  overlayGridAndComputeAvgColor(framNum, frame, grid_params, csv_file, inputVideoFile)   (:49-141)
  process_video(yolo_file, inputVideoFile, loadYoloBoxes=True, loadContours=True)        (:145-226)

Cell means / hues come from libofc's grid kernel (one launch per frame); the PNG encoding is host I/O.
A cell is saved right after its own rectangle is drawn (:104-109), i.e. with a white first row and
first column -- the same bytes as the ROI of the fully lined frame, which is what is written here.
"""
from __future__ import annotations

import argparse
import os

from . import grid as _grid
from .KmeanGrids import GRID_PARAMS, draw_grid_lines_host
from .computeOpticalFlowModule import ComputeOpticalFLow
from .drawGridsAndOutputCSV import hue_row_text
from .flow import to_device_u8


def overlayGridAndComputeAvgColor(framNum, frame, grid_params, csv_file, inputVideoFile):
    import cv2
    rows, cols = grid_params['rows'], grid_params['cols']
    height, width = frame.shape[:2]
    x_step, y_step = int(width / cols), int(height / rows)
    out = _grid.grid_cells(to_device_u8(frame), rows, cols, draw_lines=True, threshold=0, want=("avg_bgr", "avg_hue"))
    hues = out["avg_hue"][0].cpu().numpy()
    draw_grid_lines_host(frame, rows, cols)
    tm = os.path.basename(inputVideoFile).split('.')[0]
    pat = f'OutImgs/{tm}/{str(framNum)}'
    cell_idx = 0
    for y in range(rows):
        for x in range(cols):
            x1, y1 = x * x_step, y * y_step
            cell_idx += 1
            cv2.imwrite(f'{pat}/{cell_idx}.png', frame[y1:min(y1 + y_step, height), x1:min(x1 + x_step, width)])
    first = framNum <= 2
    with open(csv_file, 'w' if first else 'a', newline='') as f:
        f.write(hue_row_text(hues, first))
    return None


def process_video(yolo_bounding_box_file, inputVideoFile, loadYoloBoxes=True, loadContours=True):
    import cv2
    if loadYoloBoxes or loadContours:
        raise NotImplementedError("YOLO / contour overlays need yolo_labels.txt / Contours/ which the reference "
                                  "does not ship; run with --noyolo --nocontour")
    cap = cv2.VideoCapture(inputVideoFile)
    if not cap.isOpened():
        raise FileNotFoundError(f"cannot open video {inputVideoFile!r}")
    frameNum = 1
    ret, frame = cap.read()
    if not ret:
        raise ValueError(f"{inputVideoFile!r} has no frames")
    compflow = ComputeOpticalFLow(frame)
    tm = os.path.basename(inputVideoFile).split('.')[0]
    while cap.isOpened():
        ret, frame_rgb = cap.read()
        if not ret:
            break
        frame_optical = compflow.compute(frame_rgb)
        frameNum = frameNum + 1
        dir_path = f'OutImgs/{tm}/{str(frameNum)}'
        if not os.path.exists(dir_path):
            os.makedirs(dir_path)
        print("\n\n frameNum: ", frameNum)
        overlayGridAndComputeAvgColor(frameNum, frame_optical, GRID_PARAMS,
                                      csv_file=f"{inputVideoFile}_rgb_values.csv", inputVideoFile=inputVideoFile)
    cap.release()


def main(argv=None):
    parser = argparse.ArgumentParser(description='Example script with argparse')
    parser.add_argument('--noyolo', action='store_false', help='do not load yolo bounding boxes')
    parser.add_argument('--nocontour', action='store_false', help='do not use contour detection')
    parser.add_argument("--path", required=True, help="Path to the input video")
    args = parser.parse_args(argv)
    print('noyolo flag is set' if args.noyolo else 'noyolo flag is not set')
    process_video("yolo_labels.txt", args.path, args.noyolo, args.nocontour)


if __name__ == "__main__":
    main()
