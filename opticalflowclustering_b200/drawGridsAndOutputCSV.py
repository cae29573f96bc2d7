"""Drop-in for the reference's k-means-color-clustering/drawGridsAndOutputCSV.py grid stage.

  overlayGridAndComputeAvgColor(framNum, frame, grid_params, csv_file)       (:47-135)

Per-cell mean BGR -> uint8 (truncation) -> BGR2HSV hue, one CSV row per frame: columns
``cell_0..cell_{n-1}``, hues written as float strings (``'60.0'``), header iff
``framNum <= 2``, append otherwise (:125-135).  The cell loop of the reference is one GPU
launch here (libofc grid kernel, one CTA per cell); the white 1-px rectangles the reference
draws while looping (so that later cells see their neighbours' lines) are
reproduced by the kernel and then drawn on the caller's frame.
"""
from __future__ import annotations

import numpy as np

from . import grid as _grid
from .KmeanGrids import draw_grid_lines_host
from .flow import to_device_u8

GRID_PARAMS = {'rows': 10, 'cols': 10, 'cell_width': 10, 'cell_height': 100}     # drawGridsAndOutputCSV.py:165


def hue_row_text(hues, with_header: bool) -> str:
    """The text pandas writes for ``DataFrame([avg_hsv_colors_flat]).to_csv(index=False)`` where each
    value is ``','.join(map(str, [float hue]))`` -> ``'60.0'`` (:125-135)."""
    n = len(hues)
    head = ','.join(f"cell_{i}" for i in range(n)) + '\n' if with_header else ''
    return head + ','.join(str(float(h)) for h in hues) + '\n'


def overlayGridAndComputeAvgColor(framNum, frame, grid_params, csv_file, inputVideoFile=None):
    rows, cols = grid_params['rows'], grid_params['cols']
    dev = to_device_u8(frame)
    out = _grid.grid_cells(dev, rows, cols, draw_lines=True, threshold=0, want=("avg_bgr", "avg_hue"))
    hues = out["avg_hue"][0].cpu().numpy()
    draw_grid_lines_host(frame, rows, cols)
    first = framNum <= 2
    with open(csv_file, 'w' if first else 'a', newline='') as f:
        f.write(hue_row_text(hues, first))
    return None


def process_video(yolo_bounding_box_file, inputVideoFile, inputVideoFileExtension, loadYoloBoxes=True,
                  loadContours=True):
    """drawGridsAndOutputCSV.py:139-225 without the GUI: reads ``<video><ext>`` for the frame count
    and the pre-rendered ``<video>_optical<ext>`` for the pixels, one CSV row per frame."""
    import cv2
    if loadYoloBoxes or loadContours:
        raise NotImplementedError("YOLO / contour overlays need files the reference does not ship; "
                                  "use --noyolo --nocontour")
    cap = cv2.VideoCapture(inputVideoFile + inputVideoFileExtension)
    cap_optical_flow = cv2.VideoCapture(inputVideoFile + "_optical" + inputVideoFileExtension)
    frameNum = 1
    cap.read()
    while cap.isOpened():
        ret, _frame_rgb = cap.read()
        ret2, frame_optical = cap_optical_flow.read()
        if not ret or not ret2:
            break
        frameNum = frameNum + 1
        print("\n\n frameNum: ", frameNum)
        overlayGridAndComputeAvgColor(frameNum, frame_optical, GRID_PARAMS, csv_file="rgb_values.csv")
    cap.release()
    cap_optical_flow.release()
