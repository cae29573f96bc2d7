"""Drop-in for the reference's k-means-color-clustering/computeOpticalFlow.py CLI.

    python -m opticalflowclustering_b200.computeOpticalFlow -i video.mp4

Writes the same files: ``<in>onlyOpticalflow.mp4`` (MJPG, flow visualisation, :31,129),
``<in>_opticalFlow.csv`` (``,Frame,Average Magnitude``, :146-149) and, when matplotlib is
installed, ``<in>_squares.png`` (:152-155).  Decode/encode stay in cv2 on the host; gray,
Farneback flow, the HSV visualisation and the per-frame mean magnitude come from libofc.so,
a chunk of consecutive frames per launch sequence (ClipPipeline).
"""
from __future__ import annotations

import argparse

import numpy as np
import torch

from .pipeline import ClipPipeline


def magnitude_csv_text(means) -> str:
    """``pd.DataFrame({'Frame': x, 'Average Magnitude': y}).to_csv(path)`` (:146-149): index column,
    float32 values printed with their shortest repr."""
    lines = [",Frame,Average Magnitude"]
    for i, m in enumerate(means):
        lines.append(f"{i},{i},{np.float32(m)}")
    return "\n".join(lines) + "\n"


def run(input_path, chunk_frames: int = 17):
    import cv2 as cv
    cap = cv.VideoCapture(input_path)
    if not cap.isOpened():
        raise FileNotFoundError(f"cannot open video {input_path!r}")
    number_of_videoFrames = int(cap.get(cv.CAP_PROP_FRAME_COUNT))
    w, h = int(cap.get(3)), int(cap.get(4))
    writer = cv.VideoWriter(input_path + "onlyOpticalflow.mp4", cv.VideoWriter_fourcc(*'MJPG'),
                            cap.get(cv.CAP_PROP_FPS), (w, h))
    ret, first_frame = cap.read()
    if not ret:
        raise ValueError(f"{input_path!r} has no frames")
    print("Type1:", first_frame.dtype)
    pipe = ClipPipeline(w, h, chunk_frames=chunk_frames, draw_lines=False)
    means = []
    state = {"n": 0}

    def frames():
        yield first_frame
        while True:
            ok, frame = cap.read()
            if not ok:
                return
            yield frame

    def on_pairs(first_pair, res):
        viz = res["viz"].numpy()
        mag = res["mean_magnitude"].numpy()
        for p in range(len(mag)):
            print("Average Magnitude of optical flow ", np.float32(mag[p]))
            means.append(mag[p])
            writer.write(viz[p])
            state["n"] += 1
            print("Number of VideoFrames processed", state["n"], "/", number_of_videoFrames)

    # decode -> pinned staging -> upload on a copy stream while the previous chunk computes
    pipe.process_stream(frames(), on_pairs=on_pairs, want_viz=True)
    with open(input_path + '_opticalFlow.csv', 'w', newline='') as f:
        f.write(magnitude_csv_text(means))
    try:
        import matplotlib.pyplot as plt
        plt.plot(list(range(len(means))), means, color='black')
        plt.savefig(input_path + "_squares.png")
    except ImportError:
        print("matplotlib not installed: skipping", input_path + "_squares.png")
    cap.release()
    writer.release()
    return means


def main(argv=None):
    parser = argparse.ArgumentParser(prog='OpticalFlow', description='find optical flow of video')
    parser.add_argument('-i', '--input')
    args = parser.parse_args(argv)
    run(args.input)


if __name__ == "__main__":
    main()
