"""Standard SSD-MobileNet-v1 300x300 anchor set for the benchmarks (an INPUT of the decode; kept apart
from oracle/ so that benchmark code never imports the oracle)."""
import numpy as np


def ssd_anchors():
    out = []
    grids = [19, 10, 5, 3, 2, 1]
    scales = [0.2 + (0.95 - 0.2) * i / 5 for i in range(6)] + [1.0]
    for k, g in enumerate(grids):
        if k == 0:
            specs = [(0.1, 1.0), (scales[0], 2.0), (scales[0], 0.5)]
        else:
            specs = [(scales[k], 1.0), (scales[k], 2.0), (scales[k], 0.5), (scales[k], 3.0),
                     (scales[k], 1.0 / 3), (np.sqrt(scales[k] * scales[k + 1]), 1.0)]
        for y in range(g):
            for x in range(g):
                for s, ar in specs:
                    out.append(((y + 0.5) / g, (x + 0.5) / g, s / np.sqrt(ar), s * np.sqrt(ar)))
    return np.array(out, dtype=np.float32)
