"""Loading helpers for tests/golden/*.npz (generated from the reference by oracle/make_golden.py)."""
import hashlib
import os

import numpy as np
import torch

from deepdish_b200.scene import Scene, SceneBatch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LABELS3 = ["person", "bicycle", "car"]


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=True)


def tracker_batches(g):
    """Per-frame SceneBatch list (1 stream) for a tracker fixture: stored inputs, or regenerated from
    the Scene seed and verified against the stored checksum (None if this torch build's CPU RNG differs)."""
    if "in_feat" in g.files:
        out = []
        for f in range(int(g["frames"])):
            out.append(SceneBatch(torch.from_numpy(g["in_tlwh"][f:f + 1].copy()), torch.from_numpy(g["in_conf"][f:f + 1].copy()),
                                  torch.from_numpy(g["in_label"][f:f + 1].copy()), torch.from_numpy(g["in_feat"][f:f + 1].copy()),
                                  torch.from_numpy(g["in_count"][f:f + 1].copy())))
        return out
    kw = {k: v for k, v in g["scene_kw"]} if len(g["scene_kw"]) else {}
    sc = Scene(1, int(g["n_obj"]), int(g["dmax"]), n_labels=3, seed=int(g["seed"]), **kw)
    out = [sc.step() for _ in range(int(g["frames"]))]
    h = hashlib.sha256()
    for b in out:
        for k in ("tlwh", "conf", "label", "feat", "count"):
            h.update(np.ascontiguousarray(getattr(b, k).numpy()).tobytes())
    if h.hexdigest() != str(g["checksum"]):
        return None
    return out


def multi_batches(g):
    """Per-frame SceneBatch list (all streams) of the multi-stream fixture, regenerated from its Scene seed and
    verified against the stored checksum (None if this torch build's CPU RNG differs)."""
    sc = Scene(int(g["streams"]), int(g["n_obj"]), int(g["dmax"]), n_labels=3, seed=int(g["seed"]))
    out = [sc.step() for _ in range(int(g["frames"]))]
    h = hashlib.sha256()
    for b in out:
        for k in ("tlwh", "conf", "label", "feat", "count"):
            h.update(np.ascontiguousarray(getattr(b, k).numpy()).tobytes())
    if h.hexdigest() != str(g["checksum"]):
        return None
    return out


def yolo_full_head(g):
    """The full-size synthetic YOLOv5 head of yolo_full.npz ([frames, 25200, 85] f32), regenerated from its seed with the
    generator's own synth_yolo_head and verified against the stored checksum (None if numpy's Generator stream differs)."""
    from oracle.make_golden import synth_yolo_head
    head = synth_yolo_head(np.random.default_rng(int(g["seed"])), int(g["frames"]), int(g["na"]))
    if hashlib.sha256(head.tobytes()).hexdigest() != str(g["checksum"]):
        return None
    return head
