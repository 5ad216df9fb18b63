"""Parity at BASELINE.json's full single-GPU size (configs[2], "C3": 1024 concurrent streams x ~50 detections,
nn_budget 100, max_age 60): the scene is generated on the device like bench.py's, and

  * a sample of 32 streams is replayed through the oracle on the host with exactly the same inputs: det->track ids
    bit-exact on every tick, ids / states / hits / age / time_since_update / deleted lists / counters bit-exact and
    Kalman state within 1e-4 at the end;
  * size-independent property over ALL 1024 streams: the result does not depend on how the work is scheduled --
    3 stream chunks with the exact f32 gallery pass give bit-identical ids, track tables and Kalman state to 1 chunk
    with the half-precision pre-pass, and the chunked count reduction equals the sum over streams."""
import numpy as np
import pytest
import torch

from deepdish_b200.scene import Scene, SceneBatch
from tests.parity import OracleStreams, compare_stream, LABELS3

pytestmark = pytest.mark.gpu

S, NOBJ, DMAX, TMAX, BUDGET, MAX_AGE, TICKS = 1024, 50, 64, 128, 100, 60, 36


def test_c3_full_size_sampled_oracle_and_schedule_invariance():
    from deepdish_b200 import _lib
    from deepdish_b200.batched import BatchedTracker
    sample = list(range(0, S, 32))
    assert len(sample) == 32
    sc = Scene(S, NOBJ, DMAX, n_labels=3, seed=77, device="cuda")
    frames = [sc.step() for _ in range(TICKS)]
    a = BatchedTracker(S, LABELS3, max_tracks=TMAX, max_dets=DMAX, budget=BUDGET, max_age=MAX_AGE)
    orc = OracleStreams(len(sample), LABELS3, budget=BUDGET, max_age=MAX_AGE)
    ids_a = []
    idx = torch.as_tensor(sample, device="cuda")
    for f, b in enumerate(frames):
        got = a.step(b).cpu().numpy().copy()
        ids_a.append(got)
        hb = SceneBatch(*(getattr(b, k)[idx].cpu() for k in SceneBatch.__slots__))
        exp = orc.step(hb)
        for k, s in enumerate(sample):
            n = int(hb.count[k])
            assert list(got[s, :n]) == exp[k], (f, s)
    a.check()
    va = a.host_view()
    sub = {k: v[sample] for k, v in va.items()}
    for k in range(len(sample)):
        compare_stream(orc.trk[k], orc.cnt[k], sub, k, LABELS3)
    tot_a = a.reduce_counts().cpu().numpy().copy()
    np.testing.assert_array_equal(tot_a, va["counts"].sum(axis=0))
    assert int(va["n_tracks"].min()) > 20 and tot_a[:, 2].sum() > 0
    del a
    torch.cuda.empty_cache()
    # ---- same frames, different schedule: 3 chunks on 3 CUDA streams, exact f32 gallery pass
    b3 = BatchedTracker(S, LABELS3, max_tracks=TMAX, max_dets=DMAX, budget=BUDGET, max_age=MAX_AGE, n_chunks=3,
                        gallery_impl="exact")
    for f, b in enumerate(frames):
        b3.step(b, join=False, reduce=True)
        b3.join()
        np.testing.assert_array_equal(b3.det_track_id.cpu().numpy(), ids_a[f], err_msg="tick %d" % f)
    b3.check()
    vb = b3.host_view()
    np.testing.assert_array_equal(b3.total_counts.cpu().numpy(), tot_a)
    for name in ("n_tracks", "next_id", "n_deleted", "counts"):
        np.testing.assert_array_equal(va[name], vb[name], err_msg=name)
    live = np.arange(TMAX)[None, :] < va["n_tracks"][:, None]
    np.testing.assert_array_equal(va["order"][live], vb["order"][live])
    rows = np.repeat(np.arange(S), va["n_tracks"])
    slots = va["order"][live]
    for name in ("track_id", "state", "hits", "age", "tsu", "gal_len", "gal_pos", "mean", "cov"):
        np.testing.assert_array_equal(va[name][rows, slots], vb[name][rows, slots], err_msg=name)   # bit for bit


def test_c4_shard_size_matching_kernels_agree():
    """BASELINE configs[3] at its 8-GPU shard size (512 streams x ~180 detections, ~260 tracks): the 4-warp matching
    kernel (k_match_cta, the default at this size) and the one-warp kernel must produce bit-identical ids and track
    tables for every stream and tick -- any shared-memory race among the four warps would show up here -- and a
    sample of streams is replayed through the oracle."""
    from deepdish_b200 import _lib
    from deepdish_b200.batched import BatchedTracker
    S4, NOBJ4, D4, T4, TICKS4 = 512, 200, 224, 384, 14
    import os
    sc = Scene(S4, NOBJ4, D4, n_labels=3, seed=int(os.environ.get("DD_C4_SEED", "91")), device="cuda")
    frames = [sc.step() for _ in range(TICKS4)]
    sample = [0, 255, 511]
    orc = OracleStreams(len(sample), LABELS3, budget=BUDGET, max_age=MAX_AGE)
    idx = torch.as_tensor(sample, device="cuda")
    runs = []
    for warps in (4, 1):
        bt = BatchedTracker(S4, LABELS3, max_tracks=T4, max_dets=D4, budget=BUDGET, max_age=MAX_AGE, match_warps=warps)
        ids = []
        for f, b in enumerate(frames):
            got = bt.step(b).cpu().numpy().copy()
            ids.append(got)
            if warps == 4:
                hb = SceneBatch(*(getattr(b, k)[idx].cpu() for k in SceneBatch.__slots__))
                exp = orc.step(hb)
                for k, s in enumerate(sample):
                    assert list(got[s, :int(hb.count[k])]) == exp[k], (f, s)
        bt.check()
        runs.append((ids, bt.host_view(["n_tracks", "order", "track_id", "state", "hits", "tsu", "next_id", "mean"])))
        del bt
        torch.cuda.empty_cache()
    (ia, va), (ib, vb) = runs
    for f in range(TICKS4):
        np.testing.assert_array_equal(ia[f], ib[f], err_msg="tick %d" % f)
    np.testing.assert_array_equal(va["n_tracks"], vb["n_tracks"])
    np.testing.assert_array_equal(va["next_id"], vb["next_id"])
    live = np.arange(T4)[None, :] < va["n_tracks"][:, None]
    np.testing.assert_array_equal(va["order"][live], vb["order"][live])
    rows, slots = np.repeat(np.arange(S4), va["n_tracks"]), va["order"][live]
    for name in ("track_id", "state", "hits", "tsu", "mean"):
        np.testing.assert_array_equal(va[name][rows, slots], vb[name][rows, slots], err_msg=name)
    assert int(va["n_tracks"].max()) > 200


def test_c2_full_batch_nms_properties():
    """BASELINE configs[1] at full size (YOLOv5s head [64, 25200, 85], decode + confidence / box filter + NMS): the
    first 2 frames against the oracle, all 64 through size-independent properties of non_max_suppression
    (preprocessing.py:6-73): keep-lists in strictly descending score order, no kept box covering another kept box of
    lower score by more than max_bbox_overlap of the other's area, every dropped candidate covered by some kept box
    of higher score, and idempotence (NMS of the kept boxes keeps them all, in order)."""
    from oracle import detect as odet
    from oracle.make_golden import synth_yolo_head
    from deepdish_b200 import ops
    from tests.test_gpu_detect import COCO
    rng = np.random.default_rng(31)
    head = synth_yolo_head(rng, 64, 25200, hot=0.01)
    wanted = ["person", "bicycle", "car", "motorbike", "bus", "truck"]
    mask = torch.tensor([1 if n in wanted else 0 for n in COCO], dtype=torch.uint8, device="cuda")
    out = ops.yolo_decode(torch.from_numpy(head).cuda(), mask, 0.25, (640, 480), (640, 480), ncap=1024)
    keep, nkeep = ops.nms(out["tlwh"], out["score"], out["count"], 0.6)
    assert int(out["flags"].sum()) == 0
    tl, sc, cn = out["tlwh"].cpu().numpy(), out["score"].cpu().numpy(), out["count"].cpu().numpy()
    kp, nk = keep.cpu().numpy(), nkeep.cpu().numpy()
    for f in range(2):
        tlwh, cls, score, anchor = odet.yolo_decode(head[f], 640, 480, COCO, wanted, 0.25)
        ib, kept = odet.box_filter(list(tlwh), 640, 480)
        assert list(kp[f, :nk[f]]) == odet.non_max_suppression(ib, 0.6, score[kept])

    def covered(a, b):          # fraction of b's area (+1 convention) covered by a: the reference's overlap measure
        x1, y1 = np.maximum(a[0], b[:, 0]), np.maximum(a[1], b[:, 1])
        x2, y2 = np.minimum(a[0] + a[2], b[:, 0] + b[:, 2]), np.minimum(a[1] + a[3], b[:, 1] + b[:, 3])
        w, h = np.maximum(0, x2 - x1 + 1), np.maximum(0, y2 - y1 + 1)
        return w * h / ((b[:, 2] + 1) * (b[:, 3] + 1))

    total = 0
    for f in range(64):
        n, k = int(cn[f]), int(nk[f])
        ids = kp[f, :k]
        assert n > 100 and 0 < k <= n and len(set(ids.tolist())) == k
        s = sc[f, ids]
        assert np.all(s[:-1] > s[1:])                                   # descending pick order
        kb = tl[f, ids]
        dropped = np.setdiff1d(np.arange(n), ids)
        for i in range(k):
            assert not np.any(covered(kb[i], kb[i + 1:]) > 0.6)         # survivors are not suppressed by earlier picks
        for d in dropped[:: max(1, len(dropped) // 40)]:                # a sample of the dropped candidates
            higher = ids[sc[f, ids] > sc[f, d]]
            assert np.any(covered_by(tl[f, higher], tl[f, d]) > 0.6)
        total += k
    # idempotence on the device: NMS of the survivors keeps all of them in the same order
    D2 = int(nk.max())
    b2 = torch.zeros((64, D2, 4), dtype=torch.float64, device="cuda")
    s2 = torch.zeros((64, D2), dtype=torch.float32, device="cuda")
    for f in range(64):
        b2[f, :nk[f]] = torch.from_numpy(tl[f, kp[f, :nk[f]]]).cuda()
        s2[f, :nk[f]] = torch.from_numpy(sc[f, kp[f, :nk[f]]]).cuda()
    k2, n2 = ops.nms(b2, s2, nkeep, 0.6)
    np.testing.assert_array_equal(n2.cpu().numpy(), nk)
    k2 = k2.cpu().numpy()
    for f in range(64):
        assert list(k2[f, :nk[f]]) == list(range(nk[f]))
    assert total > 64 * 50


def covered_by(boxes, b):
    """max-over-rows helper: fraction of box b's (+1) area covered by each of `boxes`."""
    x1, y1 = np.maximum(boxes[:, 0], b[0]), np.maximum(boxes[:, 1], b[1])
    x2, y2 = np.minimum(boxes[:, 0] + boxes[:, 2], b[0] + b[2]), np.minimum(boxes[:, 1] + boxes[:, 3], b[1] + b[3])
    w, h = np.maximum(0, x2 - x1 + 1), np.maximum(0, y2 - y1 + 1)
    return w * h / ((b[2] + 1) * (b[3] + 1))
