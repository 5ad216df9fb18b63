"""numpy driver for the host emulation of the kernel bodies.  TEST SCAFFOLDING (see dd_hostemu.cpp)."""
import ctypes

import numpy as np

from deepdish_b200 import _lib as L
from . import build as _build

_emu = None


def emu():
    global _emu
    if _emu is None:
        _emu = ctypes.CDLL(_build.build())
    return _emu


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class HostEmuTracker:
    """Same call sequence as deepdish_b200.batched.BatchedTracker, on host memory."""

    def __init__(self, cfg, line=(320.0, 0.0, 320.0, 480.0)):
        self.cfg = cfg
        self.lay = L.TrackerLayout()
        assert emu().ddh_tracker_layout_query(ctypes.byref(cfg), ctypes.byref(self.lay)) == 0
        self.blob = np.zeros(self.lay.total_bytes, dtype=np.uint8)
        self.v = {}
        for name, (dt, shape) in L.field_specs(cfg).items():
            off = getattr(self.lay, name)
            n = int(np.prod(shape)) * np.dtype(dt).itemsize
            self.v[name] = self.blob[off:off + n].view(dt).reshape(shape)
        assert emu().ddh_tracker_init(_p(self.blob), ctypes.byref(cfg)) == 0
        self.line = np.asarray(line, dtype=np.float64)
        self.det_track_id = np.zeros((cfg.n_streams, cfg.max_dets), dtype=np.int32)

    def predict(self):
        assert emu().ddh_tracker_predict(_p(self.blob), ctypes.byref(self.cfg)) == 0

    def update(self, tlwh, conf, label, feat, count):
        self._keep = [np.ascontiguousarray(tlwh, np.float64), np.ascontiguousarray(conf, np.float32),
                      np.ascontiguousarray(label, np.int32), np.ascontiguousarray(feat, np.float32),
                      np.ascontiguousarray(count, np.int32)]
        a = self._keep
        rc = emu().ddh_tracker_update(_p(self.blob), ctypes.byref(self.cfg), _p(a[0]), _p(a[1]), _p(a[2]),
                                      _p(a[3]), _p(a[4]), _p(self.det_track_id))
        assert rc == 0
        return self.det_track_id

    def countline(self):
        assert emu().ddh_tracker_countline(_p(self.blob), ctypes.byref(self.cfg), _p(self.line), 0) == 0


def lsap(cost):
    cost = np.ascontiguousarray(cost, np.float64)
    nr, nc = cost.shape
    out = np.full(nr, -1, dtype=np.int32)
    rc = emu().ddh_lsap(_p(cost), nr, nc, _p(out))
    return rc, out


def set_difference_order(a, m):
    a = np.ascontiguousarray(a, np.int32)
    m = np.ascontiguousarray(m, np.int32)
    out = np.zeros(max(len(a), 1), dtype=np.int32)
    n = emu().ddh_set_difference_order(_p(a), len(a), _p(m), len(m), _p(out))
    return out[:n].tolist()


def intersection(segs):
    segs = np.ascontiguousarray(segs, np.float64).reshape(-1, 8)
    out = np.zeros(len(segs), dtype=np.int32)
    emu().ddh_intersection(_p(segs), len(segs), _p(out))
    return out
