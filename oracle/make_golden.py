"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) in the build
container.  TEST INFRASTRUCTURE -- run as ``python -m oracle.make_golden`` from the repo root.

Every fixture stores the inputs (or, for the long trajectories, the Scene seed plus a checksum of the
regenerated inputs) and the reference's outputs.  tests/test_golden.py checks the oracle against
them on any machine (the reference tree is not needed there); the GPU tests check the CUDA path
against the same files.
"""
import hashlib
import os

import numpy as np

from oracle import refload, countline as oc, detect as odet

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
LABELS3 = ["person", "bicycle", "car"]


def scene_checksum(batches):
    h = hashlib.sha256()
    for b in batches:
        for k in ("tlwh", "conf", "label", "feat", "count"):
            h.update(np.ascontiguousarray(getattr(b, k).numpy()).tobytes())
    return h.hexdigest()


def run_reference_tracker(R, batches, labels, budget, max_age, n_init=3, tcap=64, stream=0):
    """Reference Tracker + the reference's own Pipeline.process_results on one stream (`stream` of the batches)."""
    metric = R.deep_sort_nn_matching.NearestNeighborDistanceMetric("cosine", 0.2, budget)
    trk = R.deep_sort_tracker.Tracker(metric, max_iou_distance=0.7, max_age=max_age, n_init=n_init)
    cnt = refload.RefCounter(trk, labels, oc.default_line(640, 480))
    F, D = len(batches), batches[0].tlwh.shape[1]
    det_ids = np.full((F, D), -1, np.int32)
    n_tracks = np.zeros(F, np.int32)
    ids = np.full((F, tcap), -1, np.int32)
    states = np.zeros((F, tcap), np.int32)
    tsu = np.zeros((F, tcap), np.int32)
    means = np.zeros((F, tcap, 8))
    covs = np.zeros((F, tcap, 8, 8))
    deleted = np.full((F, tcap), -1, np.int32)
    counts = np.zeros((F, len(labels), 4), np.int64)
    track_labels = np.full((F, tcap), -1, np.int32)
    gal_len = np.zeros((F, tcap), np.int32)        # len(metric.samples[track]) after the update (nn_matching.py:137-154)
    for f, b in enumerate(batches):
        tlwh, conf, lab, feat = b.stream(stream)
        dets = [R.deep_sort_detection.Detection(tlwh[i], labels[lab[i]], conf[i], feat[i])
                for i in range(len(conf))]
        trk.predict()
        before = {id(t): t for t in trk.tracks}
        prev_next = trk._next_id
        trk.update(dets)
        cnt.step(dets)
        # det -> track id: a detection object is appended to track.detections on update / creation
        for t in trk.tracks + trk.deleted_tracks:
            d = t.detections[-1]
            if t.time_since_update == 0 and d in dets:
                det_ids[f, dets.index(d)] = t.track_id
        assert trk._next_id >= prev_next
        n = len(trk.tracks)
        assert n <= tcap
        n_tracks[f] = n
        for k, t in enumerate(trk.tracks):
            ids[f, k], states[f, k], tsu[f, k] = t.track_id, t.state, t.time_since_update
            means[f, k], covs[f, k] = t.mean, t.covariance
            track_labels[f, k] = labels.index(t.get_label())
            gal_len[f, k] = len(metric.samples.get(t.track_id, []))
        for k, t in enumerate(trk.deleted_tracks):
            deleted[f, k] = t.track_id
        counts[f] = cnt.counts()
    out = dict(det_ids=det_ids, n_tracks=n_tracks, ids=ids, states=states, tsu=tsu, means=means,
               covs=covs, deleted=deleted, counts=counts, track_labels=track_labels)
    if budget is None:         # the unbounded fixture also pins the gallery lengths and the final gallery of track 1
        out["gal_len"] = gal_len
        longest = max(metric.samples, key=lambda k: len(metric.samples[k]))
        out["longest_id"] = longest
        out["longest_gallery"] = np.asarray(metric.samples[longest], np.float32)
    return out


def golden_tracker(R, name, seed, n_obj, dmax, frames, budget, max_age, store_inputs, **scene_kw):
    from deepdish_b200.scene import Scene
    sc = Scene(1, n_obj, dmax, n_labels=3, seed=seed, **scene_kw)
    batches = [sc.step() for _ in range(frames)]
    out = run_reference_tracker(R, batches, LABELS3, budget, max_age)
    out.update(seed=seed, n_obj=n_obj, dmax=dmax, frames=frames, budget=budget or 0, max_age=max_age,
               checksum=scene_checksum(batches),
               scene_kw=np.array(sorted(scene_kw.items()), dtype=object) if scene_kw else np.zeros(0))
    if store_inputs:
        out.update(in_tlwh=np.stack([b.tlwh[0].numpy() for b in batches]),
                   in_conf=np.stack([b.conf[0].numpy() for b in batches]),
                   in_label=np.stack([b.label[0].numpy() for b in batches]),
                   in_feat=np.stack([b.feat[0].numpy() for b in batches]),
                   in_count=np.stack([b.count.numpy()[0] for b in batches]))
    else:
        out["means"] = out["means"][::10].copy()       # keep the file small: every 10th frame
        out["covs"] = out["covs"][-1:].copy()
    np.savez_compressed(os.path.join(OUT, name), **{k: v for k, v in out.items()}, allow_pickle=True)
    print(name, "next ids", int(out["ids"].max()), "counts", out["counts"][-1].tolist())


def golden_kalman(R):
    rng = np.random.default_rng(0)
    kf = R.deep_sort_kalman_filter.KalmanFilter()
    n, m = 40, 25
    z0 = np.c_[rng.uniform(0, 600, n), rng.uniform(0, 400, n), rng.uniform(0.2, 0.8, n), rng.uniform(40, 100, n)]
    init = [kf.initiate(z) for z in z0]
    mean = np.stack([a for a, _ in init]); cov = np.stack([b for _, b in init])
    seq_mean, seq_cov, zs = [mean.copy()], [cov.copy()], []
    for it in range(5):
        pr = [kf.predict(mean[i], cov[i]) for i in range(n)]
        mean = np.stack([a for a, _ in pr]); cov = np.stack([b for _, b in pr])
        seq_mean.append(mean.copy()); seq_cov.append(cov.copy())
        z = mean[:, :4] + rng.normal(0, 1, (n, 4)) * [2, 2, 0.01, 2]
        zs.append(z)
        up = [kf.update(mean[i], cov[i], z[i]) for i in range(n)]
        mean = np.stack([a for a, _ in up]); cov = np.stack([b for _, b in up])
        seq_mean.append(mean.copy()); seq_cov.append(cov.copy())
    proj = [kf.project(mean[i], cov[i]) for i in range(n)]
    meas = np.c_[rng.uniform(0, 600, m), rng.uniform(0, 400, m), rng.uniform(0.2, 0.8, m), rng.uniform(40, 100, m)]
    meas[:10] = mean[:10, :4] + rng.normal(0, 1.5, (10, 4)) * [1, 1, 0.01, 1]
    g4 = np.stack([kf.gating_distance(mean[i], cov[i], meas) for i in range(n)])
    g2 = np.stack([kf.gating_distance(mean[i], cov[i], meas, True) for i in range(n)])
    np.savez_compressed(os.path.join(OUT, "kalman.npz"), z0=z0, seq_mean=np.stack(seq_mean),
                        seq_cov=np.stack(seq_cov), zs=np.stack(zs), pmean=np.stack([a for a, _ in proj]),
                        pcov=np.stack([b for _, b in proj]), meas=meas, gating4=g4, gating2=g2)
    print("kalman.npz")


def golden_metric_iou(R):
    rng = np.random.default_rng(1)
    nn = R.deep_sort_nn_matching
    lens = [1, 3, 100, 17, 64]
    gal = [rng.normal(size=(l, 128)).astype(np.float32) * np.float32(rng.uniform(0.5, 2)) for l in lens]
    feats = rng.normal(size=(23, 128)).astype(np.float32)
    feats[3] = gal[2][40] + 0.01 * rng.normal(size=128).astype(np.float32)
    feats[5] = gal[4][7] + 0.05 * rng.normal(size=128).astype(np.float32)
    cos = np.stack([nn._nn_cosine_distance(g, feats) for g in gal])
    euc = np.stack([nn._nn_euclidean_distance(g, feats) for g in gal])
    n, m = 30, 21
    trk = np.c_[rng.uniform(0, 500, n), rng.uniform(0, 300, n), rng.uniform(20, 80, n), rng.uniform(40, 120, n)]
    det = np.floor(np.c_[rng.uniform(0, 500, m), rng.uniform(0, 300, m), rng.uniform(20, 80, m), rng.uniform(40, 120, m)])
    det[:10] = np.floor(trk[:10]) + 1
    iou = np.stack([R.deep_sort_iou_matching.iou(trk[i], det) for i in range(n)])
    np.savez_compressed(os.path.join(OUT, "metric_iou.npz"), gallery=np.concatenate(gal),
                        offsets=np.r_[0, np.cumsum(lens)].astype(np.int32), feats=feats, cosine=cos,
                        euclidean=euc, trk_tlwh=trk, det_tlwh=det, iou=iou)
    print("metric_iou.npz")


def golden_nms(R):
    rng = np.random.default_rng(2)
    cases = []
    for case in range(40):
        n = int(rng.integers(0, 120)) if case else 0
        k = max(1, n // 4)
        cx, cy = rng.uniform(40, 600, k), rng.uniform(40, 440, k)
        pick = rng.integers(0, k, n)
        x = np.clip(cx[pick] + rng.normal(0, 6, n), 0, 630).astype(np.int64)
        y = np.clip(cy[pick] + rng.normal(0, 6, n), 0, 470).astype(np.int64)
        w = rng.integers(10, 60, n); h = rng.integers(20, 120, n)
        boxes = np.stack([x, y, w, h], axis=1).astype(np.int64).reshape(n, 4)
        scores = (0.25 + 0.75 * (rng.permutation(n) + rng.uniform(0.1, 0.9, n)) / max(n, 1)).astype(np.float32)
        assert len(np.unique(scores)) == n
        thr = [0.6, 0.3, 0.9][case % 3]
        keep = R.deep_sort_preprocessing.non_max_suppression(boxes, thr, scores)
        cases.append((boxes, scores, thr, np.array(keep, dtype=np.int32)))
    nmax = max(len(c[0]) for c in cases)
    B = len(cases)
    boxes = np.zeros((B, nmax, 4)); scores = np.zeros((B, nmax), np.float32)
    counts = np.zeros(B, np.int32); thr = np.zeros(B); keep = np.full((B, nmax), -1, np.int32)
    nkeep = np.zeros(B, np.int32)
    for i, (b, s, t, k) in enumerate(cases):
        n = len(b)
        boxes[i, :n], scores[i, :n], counts[i], thr[i] = b, s, n, t
        keep[i, :len(k)], nkeep[i] = k, len(k)
    np.savez_compressed(os.path.join(OUT, "nms.npz"), boxes=boxes, scores=scores, counts=counts, thr=thr,
                        keep=keep, nkeep=nkeep)
    print("nms.npz", nkeep.tolist())


NPY_SCALAR_SORT = "AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL AVX512_SPR AVX2 FMA3"


def golden_nms_ties(R):
    """Tied scores at the frame sizes deepdish sees (n <= 16).  np.argsort (preprocessing.py:50) is not a stable sort
    and the order it gives equal scores depends on the numpy build and the CPU: numpy's scalar introsort -- the path
    taken on the reference's own targets (Raspberry Pi / Jetson, README.md:7-10) -- finishes every n <= 16 with an
    insertion sort, i.e. stable, so the reference picks the HIGHER index first among equal scores; numpy's AVX-512 /
    AVX2 argsort on an x86 build host orders them differently.  The fixture is therefore generated with numpy's SIMD
    dispatch switched off (NPY_DISABLE_CPU_FEATURES, re-executing this module), which pins the scalar path.
    Quantised detector heads (the default int8 YOLOv5 model multiplies two 256-level values) produce such ties.
    Also covers scores=None (rank by y2)."""
    if os.environ.get("NPY_DISABLE_CPU_FEATURES") != NPY_SCALAR_SORT:
        import subprocess
        import sys
        env = dict(os.environ, NPY_DISABLE_CPU_FEATURES=NPY_SCALAR_SORT, DD_GOLDEN_ONLY="nms_ties")
        subprocess.check_call([sys.executable, "-m", "oracle.make_golden"], env=env, cwd=ROOT)
        return
    probe = (np.arange(16) % 3).astype(np.float32)
    assert np.array_equal(np.argsort(probe), np.argsort(probe, kind="stable")), "numpy's scalar argsort is expected here"
    rng = np.random.default_rng(12)
    cases = []
    for case in range(60):
        n = int(rng.integers(2, 17))
        k = max(1, n // 3)
        cx, cy = rng.uniform(40, 600, k), rng.uniform(40, 440, k)
        pick = rng.integers(0, k, n)
        x = np.clip(cx[pick] + rng.normal(0, 5, n), 0, 630).astype(np.int64)
        y = np.clip(cy[pick] + rng.normal(0, 5, n), 0, 470).astype(np.int64)
        w = rng.integers(10, 60, n); h = rng.integers(20, 120, n)
        boxes = np.stack([x, y, w, h], axis=1).astype(np.int64).reshape(n, 4)
        levels = rng.integers(2, 6)
        scores = (rng.integers(0, levels, n) / np.float32(levels) * 0.5 + 0.25).astype(np.float32)   # few distinct values
        use_scores = case % 4 != 3
        thr = [0.6, 0.3, 0.9][case % 3]
        keep = R.deep_sort_preprocessing.non_max_suppression(boxes, thr, scores if use_scores else None)
        cases.append((boxes, scores, thr, np.array(keep, dtype=np.int32), use_scores))
    nmax, B = 16, len(cases)
    boxes = np.zeros((B, nmax, 4)); scores = np.zeros((B, nmax), np.float32)
    counts = np.zeros(B, np.int32); thr = np.zeros(B); keep = np.full((B, nmax), -1, np.int32)
    nkeep = np.zeros(B, np.int32); use = np.zeros(B, np.int32)
    for i, (b, s_, t, k_, u) in enumerate(cases):
        n = len(b)
        boxes[i, :n], scores[i, :n], counts[i], thr[i], use[i] = b, s_, n, t, int(u)
        keep[i, :len(k_)], nkeep[i] = k_, len(k_)
    np.savez_compressed(os.path.join(OUT, "nms_ties.npz"), boxes=boxes, scores=scores, counts=counts, thr=thr,
                        keep=keep, nkeep=nkeep, use_scores=use)
    print("nms_ties.npz", nkeep.tolist())


_BOXF = None


def ref_box_filter(boxes, labels, scores, w=640, h=480):
    """Reference box filter (deepdish.py:941-960 inside detect_objects) -> (int boxes [K,4] int64, kept input indices).
    Labels are replaced by the input index so that the survivors can be identified."""
    global _BOXF
    if _BOXF is None or _BOXF.p.input_size != (w, h):
        _BOXF = refload.RefBoxFilter(w, h)
    ob, ol, _ = _BOXF(list(boxes), list(range(len(boxes))), list(scores))
    return np.array(ob, dtype=np.int64).reshape(-1, 4), np.array(ol, dtype=np.int64)


def golden_box_filter(R):
    """The pre-NMS box filter as the reference runs it (deepdish.py:941-960), on adversarial float boxes: negative and
    out-of-frame corners, boxes wider than the frame, boxes above the 0.9 W H area limit, zero-size results, a NaN
    anywhere in the frame (drops every box of the frame), f32 and f64 inputs."""
    rng = np.random.default_rng(31)
    cases, res = [], {}
    for c in range(48):
        n = int(rng.integers(0, 14))
        b = np.stack([rng.uniform(-60, 700, n), rng.uniform(-60, 540, n), rng.uniform(0, 400, n), rng.uniform(0, 300, n)], 1)
        if n and c % 5 == 1:
            b[rng.integers(0, n)] = [1.5, 2.5, 637.9, 478.2]           # almost the whole viewport: rejected
        if n and c % 5 == 2:
            b[rng.integers(0, n), 2:] = [700.0, 500.0]                 # clipped to the frame, then rejected
        if n and c % 7 == 3:
            b[rng.integers(0, n), rng.integers(0, 4)] = np.nan         # one NaN drops the whole frame
        if c % 2:
            b = b.astype(np.float32)
        ib, kept = ref_box_filter([list(r) for r in b], None, [0.5] * n)
        cases.append(b.astype(np.float64))
        res["dtype%d" % c] = np.array(str(b.dtype))
        res["out%d" % c] = ib
        res["kept%d" % c] = kept
    nmax = max(len(b) for b in cases)
    boxes = np.zeros((len(cases), nmax, 4)); counts = np.zeros(len(cases), np.int32)
    for i, b in enumerate(cases):
        boxes[i, :len(b)], counts[i] = b, len(b)
    np.savez_compressed(os.path.join(OUT, "box_filter.npz"), boxes=boxes, counts=counts, **res)
    print("box_filter.npz kept", [len(res["kept%d" % c]) for c in range(len(cases))])


def synth_yolo_head(rng, frames, na, nc=80, hot=0.02):
    """[frames, na, 5+nc] f32 head in the TFLite export's format (normalised xywh, obj, cls)."""
    h = np.empty((frames, na, 5 + nc), np.float32)
    h[..., 0:2] = rng.uniform(0.05, 0.9, (frames, na, 2))
    h[..., 2:4] = rng.uniform(0.01, 0.2, (frames, na, 2))
    h[..., 4] = rng.beta(0.5, 8, (frames, na))
    h[..., 5:] = rng.beta(0.5, 4, (frames, na, nc))
    hotmask = rng.random((frames, na)) < hot          # a few confident rows
    h[..., 4][hotmask] = rng.uniform(0.5, 1.0, hotmask.sum())
    cls = rng.integers(0, 6, hotmask.sum())
    rows = np.argwhere(hotmask)
    h[rows[:, 0], rows[:, 1], 5 + cls] = rng.uniform(0.5, 1.0, len(rows))
    return h


def golden_yolo(R):
    from PIL import Image
    os.environ["DEEPDISHHOME"] = refload.REFERENCE_ROOT
    rng = np.random.default_rng(3)
    frames, na = 3, 1600
    head = synth_yolo_head(rng, frames, na)
    head[2, :, 4] *= 0.2                                   # a frame with no survivors
    head[0, 5, :4] = [0.5, 0.5, 0.99, 0.99]                # a huge box: rejected by the area filter
    head[0, 5, 4] = 0.9; head[0, 5, 5] = 0.9
    wanted = ["person", "bicycle", "car", "motorbike"]
    det = R.tools_yolov5.YOLOV5(wanted_labels=wanted, model_file="synthetic.tflite")
    names = [det.labels[i] for i in range(80)]
    img = Image.new("RGB", (640, 480))
    out = {}
    for f in range(frames):
        refload.FakeInterpreter.outputs = [head[f:f + 1]]
        boxes, labels, scores = det.detect_image(img)
        out["tlwh%d" % f] = np.array(boxes, np.float32).reshape(-1, 4)
        out["cls%d" % f] = np.array([names.index(l) for l in labels], np.int32)
        out["score%d" % f] = np.array(scores, np.float32)
        # downstream: the reference's own box filter (the loop inside Pipeline.detect_objects, deepdish.py:941-960,
        # driven by refload.RefBoxFilter) + the reference's own NMS
        ib, kept = ref_box_filter(boxes, labels, scores)
        sc = np.array(scores, np.float32)[kept]
        keep = R.deep_sort_preprocessing.non_max_suppression(np.array(ib), 0.6, sc) if len(ib) else []
        out["fbox%d" % f] = ib
        out["fidx%d" % f] = kept
        out["keep%d" % f] = np.array(keep, np.int32)
    np.savez_compressed(os.path.join(OUT, "yolo.npz"), head=head, wanted=np.array(wanted),
                        names=np.array(names), **out)
    print("yolo.npz", [len(out["tlwh%d" % f]) for f in range(frames)], [len(out["keep%d" % f]) for f in range(frames)])


def golden_yolo_full(R, frames=2, na=25200, seed=11):
    """The reference on a FULL-SIZE head (BASELINE configs[1]: 25200 anchors x 85, ~500 confident rows per frame):
    YOLOV5.detect_image + the reference's box filter + the reference's NMS.  The 17 MB head is not stored: tests
    regenerate it from the seed with this module's synth_yolo_head and verify the checksum."""
    from PIL import Image
    os.environ["DEEPDISHHOME"] = refload.REFERENCE_ROOT
    head = synth_yolo_head(np.random.default_rng(seed), frames, na)
    wanted = ["person", "bicycle", "car", "motorbike"]
    det = R.tools_yolov5.YOLOV5(wanted_labels=wanted, model_file="synthetic.tflite")
    names = [det.labels[i] for i in range(80)]
    img = Image.new("RGB", (640, 480))
    out = {}
    for f in range(frames):
        refload.FakeInterpreter.outputs = [head[f:f + 1]]
        boxes, labels, scores = det.detect_image(img)
        ib, kept = ref_box_filter(boxes, labels, scores)
        sc = np.array(scores, np.float32)[kept]
        keep = R.deep_sort_preprocessing.non_max_suppression(np.array(ib), 0.6, sc) if len(ib) else []
        out["fbox%d" % f] = np.asarray(ib, np.float64).reshape(-1, 4)
        out["fscore%d" % f] = sc
        out["fcls%d" % f] = np.array([names.index(labels[i]) for i in kept], np.int32)
        out["keep%d" % f] = np.array(keep, np.int32)
        assert len(set(sc.tolist())) == len(sc)          # unique scores: numpy's unstable sort order cannot matter
    np.savez_compressed(os.path.join(OUT, "yolo_full.npz"), frames=frames, na=na, seed=seed, wanted=np.array(wanted),
                        names=np.array(names), checksum=hashlib.sha256(head.tobytes()).hexdigest(), **out)
    print("yolo_full.npz", [len(out["fbox%d" % f]) for f in range(frames)], [len(out["keep%d" % f]) for f in range(frames)])


def golden_ssd_post(R):
    from PIL import Image
    rng = np.random.default_rng(4)
    label_path = os.path.join(refload.REFERENCE_ROOT, "detectors", "mobilenet", "labels.txt")
    names = [l.strip() for l in open(label_path)]
    wanted = ["person", "bicycle", "car", "motorcycle", "bus"]
    refload.FakeInterpreter.input_shape = (1, 300, 300, 3)
    refload.FakeInterpreter.outputs = [np.zeros((1, 10, 4), np.float32), np.zeros((1, 10), np.float32),
                                       np.zeros((1, 10), np.float32), np.zeros(1, np.float32)]
    det = R.tools_ssd_mobilenet.SSD_MOBILENET(wanted_labels=wanted, model_file="x.tflite", label_file=label_path)
    img = Image.new("RGB", (640, 480))
    cases = 60
    ob = np.zeros((cases, 10, 4), np.float32); ocl = np.zeros((cases, 10), np.float32)
    osc = np.zeros((cases, 10), np.float32)
    res = {}
    for c in range(cases):
        k = 3
        cy, cx = rng.uniform(0.2, 0.8, k), rng.uniform(0.2, 0.8, k)
        p = rng.integers(0, k, 10)
        y0 = cy[p] + rng.normal(0, 0.02, 10); x0 = cx[p] + rng.normal(0, 0.02, 10)
        hh = rng.uniform(0.1, 0.3, 10); ww = rng.uniform(0.05, 0.2, 10)
        ob[c] = np.stack([y0, x0, y0 + hh, x0 + ww], 1)
        ocl[c] = rng.choice([0, 1, 2, 3, 5, 16, 40], 10).astype(np.float32)
        osc[c] = np.sort(rng.uniform(0.2, 1.0, 10).astype(np.float32))[::-1]
        refload.FakeInterpreter.outputs = [ob[c:c + 1].copy(), ocl[c:c + 1].copy(), osc[c:c + 1].copy(),
                                           np.array([10], np.float32)]
        boxes, labels, scores = det.detect_image(img)
        res["tlwh%d" % c] = np.array(boxes, float).reshape(-1, 4)
        res["lab%d" % c] = np.array([names.index(l) for l in labels], np.int32)
        res["score%d" % c] = np.array(scores, np.float32)
        fb, fk = ref_box_filter(boxes, labels, scores)          # the reference's own box filter on its own float boxes
        res["fbox%d" % c], res["fidx%d" % c] = fb, fk
    refload.FakeInterpreter.input_shape = (1, 640, 640, 3)
    np.savez_compressed(os.path.join(OUT, "ssd_post.npz"), op_boxes=ob, op_classes=ocl, op_scores=osc,
                        names=np.array(names), wanted=np.array(wanted), **res)
    print("ssd_post.npz", sum(len(res["lab%d" % c]) for c in range(cases)))


def golden_tflite_adapter(R):
    """tools/tflite.py TFLITE.detect_image over tools/tflite_object_detector.py ObjectDetector._postprocess, both
    unmodified; the interpreter, the model metadata and the camera image are the only stand-ins."""
    import importlib
    import sys
    import types
    from PIL import Image
    if "tflite_support" not in sys.modules:
        ts = types.ModuleType("tflite_support")
        ts.metadata = types.ModuleType("tflite_support.metadata")
        sys.modules["tflite_support"] = ts
        sys.modules["tflite_support.metadata"] = ts.metadata
    tod = importlib.import_module("tools.tflite_object_detector")
    tfl = importlib.import_module("tools.tflite")
    rng = np.random.default_rng(8)
    label_list = ["person", "bicycle", "car", "motorcycle", "bus", "truck", "dog", "chair"]
    wanted = ["person", "car", "bus", "dog"]
    N, cases = 25, 48
    ob = np.zeros((cases, N, 4), np.float32); ocl = np.zeros((cases, N), np.float32)
    osc = np.zeros((cases, N), np.float32); cnt = np.zeros(cases, np.int32)
    opts, res = [], {}
    img = Image.new("RGB", (640, 480))
    for c in range(cases):
        y0 = rng.uniform(-0.05, 0.8, N); x0 = rng.uniform(-0.05, 0.8, N)
        ob[c] = np.stack([y0, x0, y0 + rng.uniform(0.02, 0.4, N), x0 + rng.uniform(0.02, 0.3, N)], 1)
        ocl[c] = rng.integers(0, len(label_list), N).astype(np.float32)
        sc = rng.uniform(0.1, 1.0, N).astype(np.float32)
        sc[rng.integers(0, N, 6)] = sc[rng.integers(0, N, 6)]          # tied scores: the sort must be stable
        osc[c] = sc
        cnt[c] = rng.integers(0, N + 1)
        o = dict(score_threshold=[0.5, 0.3, 0.0][c % 3], max_results=[-1, 5, 12][c % 3 if c % 4 else 0],
                 label_deny_list=[None, ["chair", "dog"]][c % 2], label_allow_list=[None, None, ["person", "car", "bus"]][c % 3])
        opts.append((o["score_threshold"], o["max_results"], c % 2, 1 if c % 3 == 2 else 0))
        det = object.__new__(tod.ObjectDetector)
        det._options = tod.ObjectDetectorOptions(**o)
        det._label_list = label_list
        det._input_size = (320, 320)
        det.detect = (lambda d, b, k, s, n: (lambda image: d._postprocess(b, k, s, n, image.shape[1], image.shape[0])))(
            det, ob[c], ocl[c], osc[c], int(cnt[c]))
        ad = object.__new__(tfl.TFLITE)
        ad.detector, ad.wanted_labels = det, wanted
        boxes, labels, scores = ad.detect_image(img)
        res["tlwh%d" % c] = np.array(boxes, np.int64).reshape(-1, 4)
        res["lab%d" % c] = np.array([label_list.index(l) for l in labels], np.int32)
        res["score%d" % c] = np.array(scores, np.float32)
    np.savez_compressed(os.path.join(OUT, "tflite_adapter.npz"), op_boxes=ob, op_classes=ocl, op_scores=osc, count=cnt,
                        opts=np.array(opts, np.float64), names=np.array(label_list), wanted=np.array(wanted),
                        deny=np.array(["chair", "dog"]), allow=np.array(["person", "car", "bus"]), **res)
    print("tflite_adapter.npz", sum(len(res["lab%d" % c]) for c in range(cases)))


def golden_framerecords(R):
    """deepdish/framerecords.py FrameRecords.process_boxes / process_detections / process_tracking (unmodified) driven
    like deepdish.py:993-1047 on top of the unmodified reference Tracker: an annotated object the detector misses for
    a while (process_tracking force-updates and confirms its track), an annotation overlapping a detection, an
    annotation without a detection (fed to the tracker with score 1.0) and an annotation with an unknown label."""
    import importlib
    import json
    fr = importlib.import_module("deepdish.framerecords")
    nn, Det, Trk = R.deep_sort_nn_matching, R.deep_sort_detection.Detection, R.deep_sort_tracker.Tracker
    rng = np.random.default_rng(21)
    names = {0: "person", 1: "bicycle", 2: "car"}
    F, NOBJ = 36, 6
    pos = rng.uniform([60, 60], [520, 340], (NOBJ, 2))
    vel = rng.normal(0, 3.0, (NOBJ, 2))
    size = np.stack([rng.uniform(24, 40, NOBJ), rng.uniform(50, 90, NOBJ)], 1)
    ident = rng.normal(size=(NOBJ + 1, 128))
    ident /= np.linalg.norm(ident, axis=1, keepdims=True)
    lab = rng.integers(0, 3, NOBJ)

    def feat(o, f):
        g = np.random.default_rng(1000 * f + o)
        v = ident[o] + 0.02 * g.normal(size=128)
        return (v / np.linalg.norm(v)).astype(np.float32)

    frec = fr.FrameRecords(names)
    for n in names.values():
        frec.add_annotation_label_info(n, frec.detector_labelname_to_id[n], "#000000")
    frec.add_annotation_label_info("unicorn", None, "#ffffff")
    annotations = []           # (frame, annot_track_id, label name, [x1, y1, x2, y2])
    metric = nn.NearestNeighborDistanceMetric("cosine", 0.2, 50)
    trk = Trk(metric, max_iou_distance=0.7, max_age=30, n_init=3)
    frames = []
    for f in range(F):
        pos = pos + vel
        boxes, labels, scores, owner = [], [], [], []
        for o in range(NOBJ):
            x, y = np.floor(pos[o] + rng.normal(0, 1, 2))
            w, h = np.floor(size[o] + rng.normal(0, 1, 2))
            tl = np.array([x, y, w, h], dtype=np.int64)
            missed = (o == 0 and 12 <= f < 20) or rng.random() < 0.08
            if o == 0 and f >= 8:                 # object 0 is annotated from frame 8 on (track 7); at frame 16 the
                jx, jy = (70.0, 30.0) if f >= 16 else (0.0, 0.0)      # annotation jumps: the gate fails, the track is
                annotations.append((f, 7, names[int(lab[0])],         # force-updated by process_tracking
                                    [float(x) + 0.4 + jx, float(y) + 0.3 + jy, float(x + w) + 0.2 + jx, float(y + h) + 0.1 + jy]))
            if o == 1 and 5 <= f < 15:            # a label the detector does not know, overlapping no detection
                annotations.append((f, 9, "unicorn", [600.0, 440.0, 630.0, 470.0]))   # (an overlap would be a KeyError
                                                                                      # at framerecords.py:116)
            if not missed:
                boxes.append(tl); labels.append(names[int(lab[o])]); scores.append(np.float32(0.5 + 0.45 * rng.random()))
                owner.append(o)
        if 22 <= f < 30:                          # annotated object nobody detects (track 11): enters with score 1.0
            annotations.append((f, 11, "car", [300.0 + 2 * f, 200.0, 340.0 + 2 * f, 280.0]))
        for a in [a for a in annotations if a[0] == f]:
            frec.add_annotated_track(a[0], a[1], a[2], np.array(a[3]), False, False, True, 0)
        order = np.argsort(scores)[::-1]
        boxes = [boxes[i] for i in order]; labels = [labels[i] for i in order]; scores = [scores[i] for i in order]
        owner = [owner[i] for i in order]
        b2, l2, s2 = frec.process_boxes(f, np.array(boxes).reshape(-1, 4), labels, np.array(scores))
        feats = []
        for bx in b2:                             # the encoder's stand-in: identity of the nearest object centre
            c = np.array([bx[0] + bx[2] / 2.0, bx[1] + bx[3] / 2.0])
            d2 = ((pos + size / 2.0 - c) ** 2).sum(1)
            o = int(np.argmin(d2)) if d2.min() < 40 ** 2 else NOBJ
            feats.append(feat(o, f))
        dets = [Det(bx, lb, sc, ft) for bx, lb, sc, ft in zip(b2, l2, s2, feats)]
        dets = frec.process_detections(f, dets)
        trk.predict()
        trk.update(dets)
        trk.tracks = frec.process_tracking(f, trk)
        frames.append(dict(
            boxes=[[int(v) for v in b] for b in boxes], labels=labels, scores=[float(s) for s in scores],
            out_boxes=[[float(v) for v in b] for b in b2], out_labels=list(l2), out_scores=[float(s) for s in s2],
            feats=[ft.tolist() for ft in feats],
            track_ids=[t.track_id for t in trk.tracks], states=[int(t.state) for t in trk.tracks],
            tsu=[int(t.time_since_update) for t in trk.tracks], hits=[int(t.hits) for t in trk.tracks],
            means=[t.mean.tolist() for t in trk.tracks], next_id=int(trk._next_id)))
    import gzip
    with gzip.open(os.path.join(OUT, "framerecords.json.gz"), "wt") as fh:
        json.dump(dict(names={str(k): v for k, v in names.items()}, annotations=annotations, frames=frames), fh)
    forced = sum(1 for fr_ in frames for t, s in zip(fr_["tsu"], fr_["states"]) if t == 0 and s == 2)
    print("framerecords.json.gz frames", len(frames), "tracks at end", len(frames[-1]["track_ids"]), "next_id", frames[-1]["next_id"])


def golden_patches(R):
    """tools/generate_detections.py:40-84 (the unmodified reference function, cv2.resize inside), :86-116
    (DummyImageEncoder) and :180-211 (create_box_encoder)."""
    import importlib
    gd = importlib.import_module("tools.generate_detections")
    rng = np.random.default_rng(5)
    H, W, N, NF = 240, 320, 24, 12
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    img[50:150, 100:200] = (rng.integers(0, 256, (100, 100, 1)) // 2 + np.arange(100)[None, :, None]).astype(np.uint8)
    boxes = np.stack([rng.integers(-20, W, N), rng.integers(-20, H, N), rng.integers(1, 90, N),
                      rng.integers(1, 160, N)], 1).astype(np.int64)
    boxes[0] = [W - 10, H - 10, 30, 60]     # mostly outside
    boxes[1] = [W + 60, 100, 20, 40]        # fully outside -> None
    boxes[2] = [10, 10, 1, 2]
    boxes[3] = [40, 20, 128, 256]           # taller than the frame: clipped, then 2:1-ish downscale
    boxes[4] = [5, 5, 300, 230]             # strong horizontal downscale

    def run(bs, shape):
        out = np.zeros((len(bs),) + shape + (3,), np.uint8)
        valid = np.zeros(len(bs), np.int32)
        for i, b in enumerate(bs):
            p = gd.extract_image_patch(img, b.copy(), shape)
            if p is not None:
                out[i], valid[i] = p, 1
        return out, valid

    patches, valid = run(boxes, (128, 64))
    # float boxes take the other arithmetic path of the same function (no truncation before astype(int))
    fboxes = boxes[:NF].astype(np.float64) + rng.uniform(-0.9, 0.9, (NF, 4))
    fboxes[:, 2:] = np.maximum(fboxes[:, 2:], 1.0)
    fpatches, fvalid = run(fboxes, (128, 64))
    # the reference's own arithmetic encoder: create_box_encoder('dummy') = extract_image_patch at 16x8 +
    # DummyImageEncoder; valid boxes only (a failed patch is replaced by np.random noise there)
    dpatches, dvalid = run(boxes, (16, 8))
    dboxes = boxes[dvalid == 1]
    dpatches = dpatches[dvalid == 1]
    dfeat = gd.create_box_encoder("dummy")(img, list(dboxes))
    flat = np.zeros((3, 16, 8, 3), np.uint8); flat[0] = 128; flat[1] = 7; flat[2, 3, 2] = (255, 0, 1)
    dflat = gd.DummyImageEncoder()(flat)
    np.savez_compressed(os.path.join(OUT, "patches.npz"), image=img, boxes=boxes, patches=patches, valid=valid,
                        fboxes=fboxes, fpatches=fpatches, fvalid=fvalid, dboxes=dboxes, dpatches=dpatches,
                        dfeat=dfeat, flat=flat, dflat=dflat)
    print("patches.npz valid", int(valid.sum()), "of", len(valid), "float", int(fvalid.sum()), "dummy", len(dboxes))


def synth_yolo3_heads(rng, net=128, n_hot=14, nc=80):
    """Three Keras-YOLOv3 output maps [g, g, 3*(5+nc)] f32 (raw logits) for a net x net input: background cells with
    very negative objectness / class logits and a few confident cells (clusters, so that do_nms has work to do)."""
    outs = []
    for g in (net // 32, net // 16, net // 8):
        o = np.empty((g, g, 3, 5 + nc), np.float32)
        o[..., 0:2] = rng.normal(0, 1.0, (g, g, 3, 2))
        o[..., 2:4] = rng.uniform(-0.6, 0.8, (g, g, 3, 2))
        o[..., 4] = rng.normal(-7.0, 1.0, (g, g, 3))
        o[..., 5:] = rng.normal(-6.0, 1.0, (g, g, 3, nc))
        for _ in range(max(2, n_hot * g * g // 336)):
            y, x, b = rng.integers(0, g), rng.integers(0, g), rng.integers(0, 3)
            for dy, dx in ((0, 0), (0, 1), (1, 0)) if rng.random() < 0.6 else ((0, 0),):      # neighbouring cells: overlaps
                yy, xx = min(g - 1, y + dy), min(g - 1, x + dx)
                o[yy, xx, b, 4] = rng.uniform(1.0, 5.0)
                cls = rng.choice([0, 1, 2, 3, 5, 7], 2, replace=False)
                o[yy, xx, b, 5 + cls[0]] = rng.uniform(1.0, 5.0)
                if rng.random() < 0.3:
                    o[yy, xx, b, 5 + cls[1]] = rng.uniform(0.5, 3.0)                          # a second label above threshold
        outs.append(o.reshape(g, g, 3 * (5 + nc)))
    return outs


def golden_yolo3(R):
    """tools/yolo.py (the Keras YOLOv3 adapter, UNMODIFIED): decode_netout + correct_yolo_boxes + do_nms + get_boxes and the
    tail of YOLO.detect_image (reversed order, label = argmax after NMS, transposed boxes x = box[1], y = box[0],
    :222-225) on synthetic output maps.  Stand-ins: tensorflow.keras (import only), the model (serves the maps) and
    the camera image.  float32 exp / sigmoid are numpy's: generated with numpy's SIMD dispatch switched off (the
    scalar libm path of the reference's ARM targets, see golden_nms_ties) -- stored as the fixture -- and once more
    with numpy's x86 SIMD exp, stored beside it so that the tests can COUNT the differences the exp makes."""
    import contextlib
    import importlib
    import io
    import sys
    import types
    from PIL import Image
    scalar = os.environ.get("NPY_DISABLE_CPU_FEATURES") == NPY_SCALAR_SORT
    if not scalar and os.environ.get("DD_YOLO3_PASS") != "simd":
        import subprocess
        env = dict(os.environ, DD_GOLDEN_ONLY="yolo3")
        subprocess.check_call([sys.executable, "-m", "oracle.make_golden"], env=dict(env, DD_YOLO3_PASS="simd"), cwd=ROOT)
        subprocess.check_call([sys.executable, "-m", "oracle.make_golden"],
                              env=dict(env, NPY_DISABLE_CPU_FEATURES=NPY_SCALAR_SORT), cwd=ROOT)
        return
    for name in ("tensorflow", "tensorflow.keras", "tensorflow.keras.models", "tensorflow.keras.preprocessing",
                 "tensorflow.keras.preprocessing.image"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["tensorflow.keras.models"].load_model = lambda *a, **k: None
    sys.modules["tensorflow.keras.preprocessing.image"].load_img = None
    sys.modules["tensorflow.keras.preprocessing.image"].img_to_array = None
    yolo = importlib.import_module("tools.yolo")
    rng = np.random.default_rng(41)
    frames, net = 12, 128
    det = object.__new__(yolo.YOLO)
    yolo.YOLO.__init__.__globals__["load_model"] = lambda *a, **k: None
    det.__init__(wanted_labels=["person", "bicycle", "car", "motorbike", "bus"], score_threshold=0.5)
    det.model_image_size = (net, net)
    det.height, det.width = det.model_image_size
    img = Image.new("RGB", (640, 480))
    maps = [synth_yolo3_heads(rng, net) for _ in range(frames)]
    res = {}
    for f in range(frames):
        det.model = types.SimpleNamespace(predict=lambda x, m=maps[f]: [o[None].copy() for o in m])
        with contextlib.redirect_stdout(io.StringIO()):
            boxes, labels, scores = det.detect_image(img)
        res["box%d" % f] = np.array(boxes, np.int64).reshape(-1, 4)
        res["lab%d" % f] = np.array([det.class_names.index(l) for l in labels], np.int32)
        res["score%d" % f] = np.array(scores, np.float32)
    path = os.path.join(OUT, "yolo3.npz")
    if scalar:
        prev = dict(np.load(path, allow_pickle=True)) if os.path.exists(path) else {}
        simd = {k: v for k, v in prev.items() if k.startswith("simd_")}
        np.savez_compressed(path, net=net, anchors=np.array(det.anchors, np.int32), names=np.array(det.class_names),
                            wanted=np.array(det.wanted_labels), thr=0.5,
                            **{"map%d_%d" % (f, k): maps[f][k] for f in range(frames) for k in range(3)}, **res, **simd)
        print("yolo3.npz (scalar exp)", [len(res["lab%d" % f]) for f in range(frames)])
    else:
        prev = dict(np.load(path, allow_pickle=True)) if os.path.exists(path) else {}
        prev.update({"simd_" + k: v for k, v in res.items()})
        np.savez_compressed(path, **prev)
        print("yolo3.npz (x86 SIMD exp pass)", [len(res["lab%d" % f]) for f in range(frames)])


def golden_tracker_multi(R, name="tracker_multi.npz", seed=105, streams=6, n_obj=14, dmax=20, frames=120, budget=30,
                         max_age=30):
    """Several cameras: one unmodified reference Tracker + Pipeline.process_results PER STREAM on the streams of one
    Scene -- what the batched tracker (stream chunks, captured ticks, ragged host batches) must reproduce stream by
    stream, and whose summed counters the count reduction must equal."""
    from deepdish_b200.scene import Scene
    sc = Scene(streams, n_obj, dmax, n_labels=3, seed=seed)
    batches = [sc.step() for _ in range(frames)]
    per = [run_reference_tracker(R, batches, LABELS3, budget, max_age, stream=s) for s in range(streams)]
    out = {k: np.stack([p[k] for p in per]) for k in ("det_ids", "n_tracks", "ids", "states", "tsu", "deleted", "counts",
                                                       "track_labels")}
    out["means"] = np.stack([p["means"][-1] for p in per])          # final frame only
    out["covs"] = np.stack([p["covs"][-1] for p in per])
    out.update(seed=seed, streams=streams, n_obj=n_obj, dmax=dmax, frames=frames, budget=budget, max_age=max_age,
               checksum=scene_checksum(batches))
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, "streams", streams, "next ids", out["ids"].max(axis=(1, 2)).tolist(), "counts", out["counts"][:, -1].sum(axis=0).tolist())


def golden_unbounded(R):
    """nn_budget=None -- the only way deepdish.py:515-516 ever builds its metric: galleries are never trimmed
    (nn_matching.py:137-154).  820 frames, 6 long-lived objects (no re-spawns): every track is matched ~740 times."""
    golden_tracker(R, "tracker_unbounded.npz", seed=104, n_obj=6, dmax=8, frames=820, budget=None, max_age=60,
                   store_inputs=False, respawn_prob=0.0, clutter_mean=0.5)


def main():
    os.makedirs(OUT, exist_ok=True)
    R = refload.load()
    if os.environ.get("DD_GOLDEN_ONLY") == "patches":
        golden_patches(R)
        return
    if os.environ.get("DD_GOLDEN_ONLY") == "framerecords":
        golden_framerecords(R)
        return
    if os.environ.get("DD_GOLDEN_ONLY") == "tflite":
        golden_tflite_adapter(R)
        return
    if os.environ.get("DD_GOLDEN_ONLY") == "unbounded":
        golden_unbounded(R)
        return
    if os.environ.get("DD_GOLDEN_ONLY") == "yolo3":
        golden_yolo3(R)
        return
    if os.environ.get("DD_GOLDEN_ONLY") == "box_filter":
        golden_box_filter(R)
        return
    if os.environ.get("DD_GOLDEN_ONLY") == "detect":
        golden_yolo(R)
        golden_ssd_post(R)
        return
    if os.environ.get("DD_GOLDEN_ONLY") == "nms_ties":
        golden_nms_ties(R)
        return
    if os.environ.get("DD_GOLDEN_ONLY") == "yolo_full":
        golden_yolo_full(R)
        return
    if os.environ.get("DD_GOLDEN_ONLY") == "multi":
        golden_tracker_multi(R)
        return
    golden_tflite_adapter(R)
    golden_framerecords(R)
    golden_patches(R)
    golden_kalman(R)
    golden_metric_iou(R)
    golden_nms(R)
    golden_box_filter(R)
    golden_nms_ties(R)
    golden_yolo(R)
    golden_yolo3(R)
    golden_ssd_post(R)
    golden_tracker(R, "tracker_small.npz", seed=101, n_obj=12, dmax=16, frames=100, budget=20, max_age=30,
                   store_inputs=True)
    golden_tracker(R, "tracker_c1.npz", seed=102, n_obj=20, dmax=24, frames=300, budget=100, max_age=60,
                   store_inputs=False)
    golden_tracker(R, "tracker_delcount.npz", seed=103, n_obj=10, dmax=12, frames=240, budget=100, max_age=5,
                   store_inputs=False, clutter_mean=0.0, respawn_prob=0.03)
    golden_unbounded(R)
    golden_tracker_multi(R)
    golden_yolo_full(R)


if __name__ == "__main__":
    main()
