"""The other BASELINE.json configurations as short, self-contained measurements that bench.py folds into its JSON line
(`configs`): C1 (the drop-in Tracker, one stream), C2 (YOLOv5 head decode + box filter + NMS, 64 frames), the C4 crowd
shard of one GPU and the C5 share of one GPU.  Every function runs on the current CUDA device, times with CUDA events
after warm-up, uses inputs larger than L2 (or rotates copies) and returns a plain dict.  No oracle / reference code is
used here: C1's CPU comparison is a callable handed in by bench.py's cpu_baseline leg."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"])
    except Exception:
        return 6650.0


def _timed(fn, iters, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(iters):
        fn(i)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


# ---------------------------------------------------------------------------------------------------- C2
def synth_yolo_head(frames, na, nc, gen, dev, hot=0.01):
    h = torch.empty((frames, na, 5 + nc), device=dev)
    h[..., 0:2] = 0.05 + 0.85 * torch.rand((frames, na, 2), device=dev, generator=gen)
    h[..., 2:4] = 0.01 + 0.19 * torch.rand((frames, na, 2), device=dev, generator=gen)
    h[..., 4] = torch.rand((frames, na), device=dev, generator=gen) ** 6
    h[..., 5:] = torch.rand((frames, na, nc), device=dev, generator=gen) ** 4
    m = torch.rand((frames, na), device=dev, generator=gen) < hot
    h[..., 4][m] = 0.5 + 0.5 * torch.rand(int(m.sum()), device=dev, generator=gen)
    idx = m.nonzero()
    cls = torch.randint(0, 6, (len(idx),), device=dev, generator=gen)
    h[idx[:, 0], idx[:, 1], 5 + cls] = 0.5 + 0.5 * torch.rand(len(idx), device=dev, generator=gen)
    return h


def c2(frames=64, iters=20, dev="cuda"):
    """BASELINE configs[1]: YOLOv5s-shaped raw head [64, 25200, 85] f32 -> decode + confidence / wanted-label filter +
    box filter + NMS.  Algorithmic bytes = the head (8.568 MB per frame)."""
    from deepdish_b200 import ops
    gen = torch.Generator(device=dev).manual_seed(0)
    NA, NC = 25200, 80
    mask = torch.zeros(NC, dtype=torch.uint8, device=dev)
    mask[:8] = 1
    heads = [synth_yolo_head(frames, NA, NC, gen, dev) for _ in range(3)]       # 3 x 548 MB, rotated: never L2-resident
    out = {}

    def dec(i):
        out["y"] = ops.yolo_decode(heads[i % 3], mask, 0.25, (640, 480), (640, 480), ncap=1024, out=out.get("y"))

    def both(i):
        dec(i)
        out["k"] = ops.nms(out["y"]["tlwh"], out["y"]["score"], out["y"]["count"], 0.6, out=out.get("k"))

    ms_dec = _timed(dec, iters)
    ms = _timed(both, iters)
    y = out["y"]
    ms_nms = _timed(lambda i: ops.nms(y["tlwh"], y["score"], y["count"], 0.6, out=out["k"]), iters)
    nbytes = frames * NA * (5 + NC) * 4
    pk = hbm_peak()
    assert int(y["flags"].max()) == 0
    # steady-state stream of batches: the NMS of batch k (latency-bound, one CTA per frame) runs on a second CUDA
    # stream under the HBM-bound decode of batch k + 1; outputs are double-buffered
    sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
    outs, dec_done, nms_done = [{}, {}], [None, None], [None, None]

    def piped(i):
        k = i & 1
        with torch.cuda.stream(sA):
            if nms_done[k] is not None:
                sA.wait_event(nms_done[k])
            outs[k]["y"] = ops.yolo_decode(heads[i % 3], mask, 0.25, (640, 480), (640, 480), ncap=1024, out=outs[k].get("y"))
            dec_done[k] = sA.record_event()
        with torch.cuda.stream(sB):
            sB.wait_event(dec_done[k])
            outs[k]["k"] = ops.nms(outs[k]["y"]["tlwh"], outs[k]["y"]["score"], outs[k]["y"]["count"], 0.6, out=outs[k].get("k"))
            nms_done[k] = sB.record_event()

    for i in range(4):
        piped(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    sA.wait_event(s); sB.wait_event(s)
    for i in range(iters):
        piped(i)
    cur = torch.cuda.current_stream()
    cur.wait_event(dec_done[0]); cur.wait_event(dec_done[1]); cur.wait_event(nms_done[0]); cur.wait_event(nms_done[1])
    e.record()
    torch.cuda.synchronize()
    ms_serial, ms = ms, s.elapsed_time(e) / iters
    return {"workload": "C2: YOLOv5s raw head [%d,25200,85] f32 -> decode + conf filter + box filter + NMS, batches streamed "
                        "(NMS of batch k under the decode of batch k+1)" % frames,
            "metric": "frames/s", "value": frames / ms * 1e3, "ms_per_batch": ms, "ms_per_batch_serial": ms_serial,
            "ms_decode": ms_dec, "ms_nms": ms_nms,
            "candidates_per_frame": float(y["count"].float().mean()), "kept_per_frame": float(out["k"][1].float().mean()),
            "roofline": {"bound": "hbm", "achieved": nbytes / ms / 1e6, "peak": pk, "unit": "GB/s",
                         "frac": nbytes / ms / 1e6 / pk, "bytes_per_batch": nbytes,
                         "serial_frac": nbytes / ms_serial / 1e6 / pk, "decode_only_frac": nbytes / ms_dec / 1e6 / pk}}


# ---------------------------------------------------------------------------------------------------- C4
def c4(streams=512, steps=12, warmup=3, preroll=110, dev="cuda", **knobs):
    """BASELINE configs[3] at its 8-GPU shard size: 512 streams x ~200 detections per frame, up to ~270 tracks per
    stream, nn_budget 100.  One tick = predict + update + count-line + count reduce for all streams of this GPU."""
    from deepdish_b200.batched import BatchedTracker
    from deepdish_b200.scene import Scene
    labels = ["person", "bicycle", "car"]
    NOBJ, DMAX, TMAX = 200, 224, 384
    bt = BatchedTracker(streams, labels, max_tracks=TMAX, max_dets=DMAX, budget=100, max_age=60, device=dev, n_chunks=2, **knobs)
    scene = Scene(streams, NOBJ, DMAX, n_labels=3, seed=77, device=dev)
    for _ in range(preroll):
        bt.step(scene.step())
    frames = [scene.step() for _ in range(warmup + steps)]
    for b in frames[:warmup]:
        bt.step(b, join=False, reduce=True)
    bt.join()
    G = float(bt.gallery_vectors().sum())
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for b in frames[warmup:]:
        bt.step(b, join=False, reduce=True)
    bt.join()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    bt.check()
    dets = sum(int(b.count.sum()) for b in frames[warmup:]) / steps
    conf = float((bt.v["state"] == 2).sum())
    tick_bytes = 512.0 * (G + 2 * dets) + 1152.0 * conf * 1.1 + 44.0 * dets           # SURVEY 8d B_trk, summed over streams
    pk = hbm_peak()
    out = {"workload": "C4 crowd shard: %d streams/GPU x ~%d dets/frame, up to %d tracks/stream, nn_budget 100" % (streams, NOBJ, TMAX),
           "metric": "tracked stream-frames/s", "value": streams / ms * 1e3, "ms_per_tick": ms,
           "gallery_rows_per_stream": G / streams, "dets_per_frame": dets / streams, "tracker_state_bytes": bt.memory_bytes(),
           "roofline": {"bound": "hbm", "achieved": tick_bytes / ms / 1e6, "peak": pk, "unit": "GB/s",
                        "frac": tick_bytes / ms / 1e6 / pk, "bytes_per_tick": tick_bytes,
                        "note": "SURVEY 8d algorithmic bytes of the whole tick (f32 gallery rows) / tick time"}}
    del bt
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------------- C5
def c5(n_yolo=1024, n_ssd=1024, steps=8, warmup=3, preroll=40, objects=40, dev="cuda"):
    """BASELINE configs[4], one GPU's share (16384 mixed streams over 8 GPUs = 1024 YOLOv5 + 1024 SSD-MobileNet streams per
    GPU): head decode + box filter + NMS + gather + tracker tick + count-line + count reduce, heads resident in HBM."""
    from deepdish_b200.batched import BatchedTracker
    from deepdish_b200.pipeline import DetectTrackPipeline, YoloFrontEnd, SsdFrontEnd
    from oracle_free_anchors import ssd_anchors
    labels = ["person", "bicycle", "car", "bus"]
    gen = torch.Generator(device=dev).manual_seed(5)
    SY, SS, D = n_yolo, n_ssd, 64
    S = SY + SS
    NA, NC = 25200, 80
    coco = ["person", "bicycle", "car", "motorbike", "aeroplane", "bus"] + ["c%02d" % i for i in range(6, 80)]
    ssd_names = ["???"] + ["c%02d" % i for i in range(1, 91)]
    ssd_names[1], ssd_names[2], ssd_names[3], ssd_names[6] = "person", "bicycle", "car", "bus"
    head = torch.empty((SY, NA, 5 + NC), device=dev)
    for lo in range(0, SY, 64):
        n = min(64, SY - lo)
        h = head[lo:lo + n]
        h[..., 0:2] = 0.1 + 0.8 * torch.rand((n, NA, 2), device=dev, generator=gen)
        h[..., 2:4] = 0.02 + 0.08 * torch.rand((n, NA, 2), device=dev, generator=gen)
        h[..., 4] = 0.2 * torch.rand((n, NA), device=dev, generator=gen)
        h[..., 5:] = 0.5 * torch.rand((n, NA, NC), device=dev, generator=gen)
        rows = torch.rand((n, NA), device=dev, generator=gen).argsort(dim=1)[:, :objects]
        fi = torch.arange(n, device=dev)[:, None].expand_as(rows)
        score = 0.5 + 0.49 * (torch.rand((n, objects), device=dev, generator=gen).argsort(dim=1).float() + 0.5) / objects
        h[fi, rows, 4] = score
        h[fi, rows, 5:] = 0.01
        cls = torch.tensor([0, 1, 2, 5], device=dev)[torch.randint(0, 4, (n, objects), device=dev, generator=gen)]
        h[fi, rows, 5 + cls] = 1.0
    rb = 0.3 * torch.randn((SS, 1917, 4), device=dev, generator=gen)
    sc = 0.3 * torch.rand((SS, 1917, 91), device=dev, generator=gen) ** 4
    hot = torch.rand((SS, 1917), device=dev, generator=gen).argsort(dim=1)[:, :12]
    fi = torch.arange(SS, device=dev)[:, None].expand_as(hot)
    sc[fi, hot, 1 + torch.tensor([0, 1, 2, 5], device=dev)[torch.randint(0, 4, (SS, 12), device=dev, generator=gen)]] = \
        0.55 + 0.44 * torch.rand((SS, 12), device=dev, generator=gen)
    anchors = torch.from_numpy(ssd_anchors()).to(dev)
    ident = torch.randn((S, D, 128), device=dev, generator=gen)
    ident = ident / ident.norm(dim=-1, keepdim=True)
    bt = BatchedTracker(S, labels, max_tracks=128, max_dets=D, budget=100, max_age=60, n_chunks=2, device=dev)
    pipe = DetectTrackPipeline(bt, [YoloFrontEnd(0, SY, coco, labels, labels, ncap=1024),
                                    SsdFrontEnd(SY, S, ssd_names, labels, labels, anchors)])

    def feats():
        f = ident + 0.02 * torch.randn((S, D, 128), device=dev, generator=gen)
        return f / f.norm(dim=-1, keepdim=True)

    heads = [head, (rb, sc)]
    fl = [feats() for _ in range(4)]
    for t in range(preroll + warmup):
        pipe.step(heads, fl[t % 4], join=False)
    bt.join()
    pipe.check()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    for t in range(steps):
        pipe.detect(heads)
    ev[1].record()
    ev[2].record()
    for t in range(steps):
        pipe.step(heads, fl[t % 4], join=False)
    bt.join()
    ev[3].record()
    torch.cuda.synchronize()
    pipe.check()
    det_ms, all_ms = ev[0].elapsed_time(ev[1]) / steps, ev[2].elapsed_time(ev[3]) / steps
    head_bytes = SY * NA * (5 + NC) * 4 + SS * 1917 * 95 * 4
    G = float(bt.gallery_vectors().sum())
    dets = float(pipe.det_count.float().sum())
    tick_bytes = head_bytes + 512.0 * (G + 2 * dets) + 1152.0 * float((bt.v["state"] == 2).sum()) * 1.1 + 44.0 * dets
    pk = hbm_peak()
    out = {"workload": "C5 share of one GPU: %d YOLOv5 (25200x85 f32) + %d SSD-MobileNet (1917x91) streams, decode + box "
                       "filter + NMS + gather + tracker tick + count-line + count reduce" % (SY, SS),
           "metric": "tracked stream-frames/s", "value": S / all_ms * 1e3, "ms_per_tick": all_ms, "detect_only_ms": det_ms,
           "dets_per_stream": dets / S, "tracks_per_stream": float(bt.v["n_tracks"].float().mean()),
           "roofline": {"bound": "hbm", "achieved": tick_bytes / all_ms / 1e6, "peak": pk, "unit": "GB/s",
                        "frac": tick_bytes / all_ms / 1e6 / pk, "bytes_per_tick": tick_bytes,
                        "detect_front_end_frac": head_bytes / det_ms / 1e6 / pk}}
    del bt, pipe, head, rb, sc
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------------- C1
def c1(frames=300, n_obj=20, dmax=24, budget=100, max_age=60, cpu_port=None):
    """BASELINE configs[0]: DeepSORT on one synthetic 640x480 stream, 300 frames, <= 20 detections per frame, 128-d
    features, nn_budget 100 -- through the drop-in ``deep_sort`` API (Detection objects in, Tracker.predict() /
    update(), Track objects out) on the GPU.  cpu_port (optional, supplied by bench.py's cpu_baseline leg): a callable
    run(batches, labels, budget, max_age) -> (seconds, final track ids, confirmed-track visits) timing the CPU port of
    the reference on the same frames on one host core."""
    from deepdish_b200.deep_sort import nn_matching
    from deepdish_b200.deep_sort.detection import Detection
    from deepdish_b200.deep_sort.tracker import Tracker
    from deepdish_b200.scene import Scene
    labels = ["person", "bicycle", "car"]
    sc = Scene(1, n_obj, dmax, n_labels=3, seed=102)
    batches = [sc.step().stream(0) for _ in range(frames)]
    trk = Tracker(nn_matching.NearestNeighborDistanceMetric("cosine", 0.2, budget), max_iou_distance=0.7, max_age=max_age, n_init=3)

    def run(tracker, mk):
        t0 = time.perf_counter()
        n_conf = 0
        for tlwh, conf, lab, feat in batches:
            dets = [mk(tlwh[i], labels[lab[i]], conf[i], feat[i]) for i in range(len(conf))]
            tracker.predict()
            tracker.update(dets)
            n_conf += sum(1 for t in tracker.tracks if t.is_confirmed() and t.time_since_update <= 1)   # deepdish.py:1053
        return time.perf_counter() - t0, n_conf

    run(trk, Detection)                                   # warm-up pass (allocations, first launches)
    trk = Tracker(nn_matching.NearestNeighborDistanceMetric("cosine", 0.2, budget), max_iou_distance=0.7, max_age=max_age, n_init=3)
    torch.cuda.synchronize()
    gpu_s, gpu_conf = run(trk, Detection)
    torch.cuda.synchronize()
    ids = [t.track_id for t in trk.tracks]
    out = {"workload": "C1: drop-in deep_sort Tracker, 1 stream 640x480, %d frames, <= %d dets/frame, nn_budget %d" % (frames, n_obj, budget),
           "metric": "frames/s", "value": frames / gpu_s, "ms_per_frame": 1e3 * gpu_s / frames,
           "note": "latency-bound at S = 1: one predict() + update() per frame incl. the host<->device copies of the "
                   "Detection list and the lazy Track views (read every frame here, like deepdish.py:1053); the batched "
                   "path is what the GPU is for"}
    if cpu_port is not None:
        cpu_s, cpu_ids, cpu_conf = cpu_port(batches, labels, budget, max_age)
        out.update(cpu_port_ms_per_frame=1e3 * cpu_s / frames, cpu_port_frames_per_s=frames / cpu_s,
                   same_final_ids_as_cpu_port=bool(ids == cpu_ids and gpu_conf == cpu_conf))
    return out


def run_all(knobs=None, cpu_port=None):
    out = {}
    for name, fn in (("c1", lambda: c1(cpu_port=cpu_port)), ("c2", c2), ("c4", lambda: c4(**(knobs or {}))), ("c5", c5)):
        try:
            out[name] = fn()
        except Exception as e:                            # a failed extra must not lose the headline line
            out[name] = {"error": repr(e)}
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c2", "c4", "c5"]
    for w in which:
        print(json.dumps({w: {"c1": c1, "c2": c2, "c4": c4, "c5": c5}[w]()}), flush=True)
