"""BatchedTracker -- S independent DeepSORT trackers stepped together on one B200.

The batched extension of the reference's ``deep_sort.tracker.Tracker``: same
``predict()`` / ``update(detections)`` call sequence (tracker.py:51-93), plus the count-line step of
``Pipeline.process_results`` (deepdish.py:1035-1114).  All state lives in one caller-owned device
blob whose layout is defined by the C ABI (include/deepdish_b200.h); this class only allocates it
with torch, passes raw pointers + the current CUDA stream to libdeepdish_b200.so, and exposes typed
torch views for inspection.  There is no CPU path.
"""
import ctypes

import torch

from . import _lib

_DT = {"int32": torch.int32, "int64": torch.int64, "float32": torch.float32,
       "float64": torch.float64}


class BatchedTracker:
    def __init__(self, n_streams, labels, max_tracks=128, max_dets=64, budget=100,
                 max_cosine_distance=0.2, max_iou_distance=0.7, max_age=30, n_init=3,
                 line=None, frame_size=(640, 480), device="cuda"):
        self.lib = _lib.lib()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("BatchedTracker runs on a CUDA device only (no CPU fallback)")
        self.labels = list(labels)
        self.cfg = _lib.make_config(n_streams, max_tracks, max_dets, budget, self.labels,
                                    max_age=max_age, n_init=n_init,
                                    max_cosine_distance=max_cosine_distance,
                                    max_iou_distance=max_iou_distance)
        self.lay = _lib.TrackerLayout()
        _lib.check(self.lib.dd_tracker_layout_query(ctypes.byref(self.cfg), ctypes.byref(self.lay)),
                   "dd_tracker_layout_query")
        self.blob = torch.empty(self.lay.total_bytes, dtype=torch.uint8, device=self.device)
        self.v = {}
        for name, (dt, shape) in _lib.field_specs(self.cfg).items():
            off = getattr(self.lay, name)
            n = 1
            for d in shape:
                n *= d
            n *= torch.empty((), dtype=_DT[dt]).element_size()
            self.v[name] = self.blob[off:off + n].view(_DT[dt]).view(shape)
        if line is None:                      # deepdish.py:739-744
            w, h = frame_size
            line = (float(int(w / 2)), 0.0, float(int(w / 2)), float(int(h)))
        lt = torch.as_tensor(line, dtype=torch.float64)
        self.line_per_stream = 1 if lt.dim() == 2 else 0
        self.line = lt.to(self.device).contiguous()
        S, D = n_streams, max_dets
        self.det_track_id = torch.full((S, D), -1, dtype=torch.int32, device=self.device)
        self.total_counts = torch.zeros((len(self.labels), 4), dtype=torch.int64, device=self.device)
        self._cfgp = ctypes.byref(self.cfg)
        self._state = self.blob.data_ptr()
        _lib.check(self.lib.dd_tracker_init(self._state, self._cfgp, self._stream()), "dd_tracker_init")

    # ------------------------------------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def predict(self):
        """Tracker.predict for every stream (tracker.py:51-57)."""
        _lib.check(self.lib.dd_tracker_predict(self._state, self._cfgp, self._stream()),
                   "dd_tracker_predict")

    def update(self, tlwh, conf, label, feat, count):
        """Tracker.update for every stream (tracker.py:59-93).

        tlwh f64 [S,Dmax,4], conf f32 [S,Dmax], label i32 [S,Dmax], feat f32 [S,Dmax,128], count i32 [S]
        -- device tensors, padded.  Returns det_track_id i32 [S,Dmax] (device; -1 = padding)."""
        S, D = self.cfg.n_streams, self.cfg.max_dets
        for t, dt, shape in ((tlwh, torch.float64, (S, D, 4)), (conf, torch.float32, (S, D)),
                             (label, torch.int32, (S, D)), (feat, torch.float32, (S, D, 128)),
                             (count, torch.int32, (S,))):
            if t.dtype != dt or tuple(t.shape) != shape or not t.is_cuda or not t.is_contiguous():
                raise ValueError("update(): expected contiguous CUDA %s %s, got %s %s on %s"
                                 % (dt, shape, t.dtype, tuple(t.shape), t.device))
        _lib.check(self.lib.dd_tracker_update(
            self._state, self._cfgp, tlwh.data_ptr(), conf.data_ptr(), label.data_ptr(),
            feat.data_ptr(), count.data_ptr(), self.det_track_id.data_ptr(), self._stream()),
            "dd_tracker_update")
        return self.det_track_id

    def update_profiled(self, tlwh, conf, label, feat, count, events5):
        """update() that records 5 CUDA events (see dd_tracker_update_profiled); events5 = ctypes
        array of handles from new_events()."""
        _lib.check(self.lib.dd_tracker_update_profiled(
            self._state, self._cfgp, tlwh.data_ptr(), conf.data_ptr(), label.data_ptr(),
            feat.data_ptr(), count.data_ptr(), self.det_track_id.data_ptr(), self._stream(), events5),
            "dd_tracker_update_profiled")
        return self.det_track_id

    def new_events(self, n=5):
        arr = (ctypes.c_void_p * n)()
        for i in range(n):
            h = ctypes.c_void_p()
            _lib.check(self.lib.dd_event_create(ctypes.byref(h)), "dd_event_create")
            arr[i] = h
        return arr

    def elapsed_ms(self, start, end):
        ms = ctypes.c_float(0)
        _lib.check(self.lib.dd_event_elapsed_ms(start, end, ctypes.byref(ms)), "dd_event_elapsed_ms")
        return ms.value

    def countline(self):
        """Count-line step (deepdish.py:1041-1112) on the state left by update()."""
        _lib.check(self.lib.dd_tracker_countline(self._state, self._cfgp, self.line.data_ptr(),
                                                 self.line_per_stream, self._stream()),
                   "dd_tracker_countline")

    def step(self, batch):
        """predict + update + countline for one SceneBatch-like object (device tensors)."""
        self.predict()
        ids = self.update(batch.tlwh, batch.conf, batch.label, batch.feat, batch.count)
        self.countline()
        return ids

    def step_host(self, host_batch):
        """End-to-end entry: pinned HOST batch -> H2D -> tick -> device-side count reduction.
        Returns (det_track_id, total_counts) device tensors; the caller reads them back."""
        dev = host_batch.to(self.device, non_blocking=True)
        ids = self.step(dev)
        return ids, self.reduce_counts()

    def reduce_counts(self):
        """Sum the per-stream counters -> i64 [C,4] (pos, neg, int, del per label), this GPU only."""
        _lib.check(self.lib.dd_tracker_count_reduce(self._state, self._cfgp,
                                                    self.total_counts.data_ptr(), self._stream()),
                   "dd_tracker_count_reduce")
        return self.total_counts

    def all_reduce_counts(self):
        """reduce_counts() followed by the NCCL all-reduce over ranks (the only collective)."""
        t = self.reduce_counts()
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    def status(self):
        flags = ctypes.c_int32(0)
        _lib.check(self.lib.dd_tracker_status(self._state, self._cfgp, ctypes.byref(flags),
                                              self._stream()), "dd_tracker_status")
        return flags.value

    def check(self):
        """Raise if any stream overflowed its capacities (never silently truncated)."""
        f = self.status()
        if f & _lib.FLAG_TRACK_OVERFLOW:
            raise RuntimeError("track capacity exceeded (max_tracks=%d)" % self.cfg.max_tracks)
        if f & _lib.FLAG_DET_OVERFLOW:
            raise RuntimeError("detection capacity exceeded (max_dets=%d)" % self.cfg.max_dets)
        if f & _lib.FLAG_LSAP_INFEASIBLE:
            raise ValueError("cost matrix is infeasible")

    # ------------------------------------------------------------------------------------------
    def host_view(self, names=None, streams=None):
        """numpy copies of state arrays (optionally a subset of streams) for inspection / tests."""
        names = names or [n for n in self.v if n not in ("gal", "cost", "gate", "det_featn")]
        out = {}
        for n in names:
            t = self.v[n]
            if streams is not None:
                t = t[torch.as_tensor(streams, device=self.device)]
            out[n] = t.cpu().numpy()
        return out

    def gallery_vectors(self):
        """Total gallery vectors of confirmed live tracks (G in SURVEY.md section 8d), per stream."""
        S, T = self.cfg.n_streams, self.cfg.max_tracks
        valid = torch.arange(T, device=self.device)[None, :] < self.v["n_tracks"][:, None]
        rows = torch.arange(S, device=self.device)[:, None].expand(S, T)[valid]
        slots = self.v["order"].long()[valid]
        live = torch.zeros((S, T), dtype=torch.bool, device=self.device)
        live[rows, slots] = True
        conf = live & (self.v["state"] == 2)
        return (self.v["gal_len"] * conf).sum(dim=1)
