"""GPU parity against fixtures produced by the UNMODIFIED reference (tests/golden/, oracle/make_golden.py):
the CUDA path is compared with the reference's own outputs, not only with the oracle."""
import numpy as np
import pytest
import torch

from tests import goldens
from tests.goldens import LABELS3

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["tracker_small.npz", "tracker_c1.npz", "tracker_delcount.npz", "tracker_unbounded.npz"])
def test_tracker_vs_reference_fixture(name):
    from deepdish_b200.batched import BatchedTracker
    g = goldens.load(name)
    batches = goldens.tracker_batches(g)
    if batches is None:
        pytest.skip("regenerated inputs do not match the fixture checksum")
    D = int(g["dmax"])
    # budget 0 in a fixture = nn_budget None, the reference's own configuration (deepdish.py:515-516); the small pool
    # and page table make the run grow both several times (pool segments attached, page table re-laid out)
    unb = int(g["budget"]) == 0
    kw = dict(pool_pages=64, seg_pages=64, page_cap=4) if unb else {}
    bt = BatchedTracker(1, LABELS3, max_tracks=96, max_dets=D, budget=int(g["budget"]) or None, max_age=int(g["max_age"]), **kw)
    stored = "in_feat" in g.files
    for f, b in enumerate(batches):
        got = bt.step(b.to("cuda")).cpu().numpy()[0]
        n = int(b.count[0])
        assert list(got[:n]) == list(g["det_ids"][f, :n]), f
        if unb:
            bt.maintain(wait=True)
        v = bt.host_view(["n_tracks", "order", "track_id", "state", "tsu", "n_deleted", "deleted", "mean", "cov", "counts", "gal_len"])
        nt = int(v["n_tracks"][0]); sl = v["order"][0, :nt]
        assert nt == int(g["n_tracks"][f])
        assert list(v["track_id"][0, sl]) == list(g["ids"][f, :nt])
        assert list(v["state"][0, sl]) == list(g["states"][f, :nt])
        assert list(v["tsu"][0, sl]) == list(g["tsu"][f, :nt])
        dl = v["deleted"][0, :int(v["n_deleted"][0])]
        assert list(v["track_id"][0, dl]) == [x for x in g["deleted"][f] if x >= 0]
        np.testing.assert_array_equal(v["counts"][0], g["counts"][f])
        if unb:
            conf = v["state"][0, sl] == 2          # metric.samples holds confirmed tracks only
            assert list(v["gal_len"][0, sl][conf]) == list(g["gal_len"][f, :nt][conf]), f
        if stored:
            np.testing.assert_allclose(v["mean"][0, sl], g["means"][f, :nt], rtol=1e-4, atol=1e-9)
            np.testing.assert_allclose(v["cov"][0, sl], g["covs"][f, :nt], rtol=1e-4, atol=1e-12)
        elif f % 10 == 0:
            np.testing.assert_allclose(v["mean"][0, sl], g["means"][f // 10, :nt], rtol=1e-4, atol=1e-9)
    bt.check()
    if name == "tracker_delcount.npz":
        assert int(v["counts"][0, :, 3].sum()) > 0       # the del counter was exercised
    if unb:
        slot = int(sl[list(v["track_id"][0, sl]).index(int(g["longest_id"]))])
        ref = g["longest_gallery"]
        assert len(ref) >= 700 and len(bt.chunks[0].segs) > 1 and bt.chunks[0].v["ptab"].shape[2] > 4
        unit = ref / np.sqrt(np.sum(ref * ref, axis=1, dtype=np.float32))[:, None]
        np.testing.assert_allclose(bt.gallery(0, slot).cpu().numpy(), unit, rtol=2e-6, atol=1e-7)


def test_drop_in_tracker_with_nn_budget_none_vs_reference_fixture():
    """The reference's own configuration -- NearestNeighborDistanceMetric("cosine", thr, None), deepdish.py:515-517 --
    through the deep_sort API mirror: 820 frames, four tracks matched > 700 times, ids / states / gallery lengths
    against the unmodified reference's run."""
    from deepdish_b200.deep_sort import nn_matching
    from deepdish_b200.deep_sort.detection import Detection
    from deepdish_b200.deep_sort.tracker import Tracker
    g = goldens.load("tracker_unbounded.npz")
    batches = goldens.tracker_batches(g)
    if batches is None:
        pytest.skip("regenerated inputs do not match the fixture checksum")
    metric = nn_matching.NearestNeighborDistanceMetric("cosine", 0.2, None)
    trk = Tracker(metric, max_iou_distance=0.7, max_age=int(g["max_age"]), n_init=3)
    for f, b in enumerate(batches):
        tlwh, conf, lab, feat = b.stream(0)
        dets = [Detection(tlwh[i], LABELS3[lab[i]], conf[i], feat[i]) for i in range(len(conf))]
        trk.predict()
        trk.update(dets)
        n = int(g["n_tracks"][f])
        assert [t.track_id for t in trk.tracks] == list(g["ids"][f, :n]), f
        assert [t.state for t in trk.tracks] == list(g["states"][f, :n]), f
        assert [t.track_id for t in trk.deleted_tracks] == [x for x in g["deleted"][f] if x >= 0]
        if f % 10 == 0:
            np.testing.assert_allclose(np.stack([t.mean for t in trk.tracks]), g["means"][f // 10, :n], rtol=1e-4, atol=1e-9)
            assert [LABELS3.index(t.get_label()) for t in trk.tracks] == list(g["track_labels"][f, :n])
    ref = g["longest_gallery"]
    got = np.asarray(metric.samples[int(g["longest_id"])])
    assert got.shape == ref.shape and len(ref) >= 700
    np.testing.assert_allclose(got, ref / np.sqrt(np.sum(ref * ref, axis=1, dtype=np.float32))[:, None], rtol=2e-6, atol=1e-7)
    assert {k: len(v) for k, v in metric.samples.items()} == {
        int(i): int(l) for i, l, s in zip(g["ids"][-1], g["gal_len"][-1], g["states"][-1]) if s == 2}


def test_nms_vs_reference_fixture():
    from deepdish_b200 import ops
    g = goldens.load("nms.npz")
    B = len(g["counts"])
    for thr in np.unique(g["thr"]):
        sel = np.nonzero(g["thr"] == thr)[0]
        keep, nkeep = ops.nms(ops._dev(g["boxes"][sel], torch.float64), ops._dev(g["scores"][sel], torch.float32),
                              ops._dev(g["counts"][sel], torch.int32), float(thr))
        keep, nkeep = keep.cpu().numpy(), nkeep.cpu().numpy()
        for k, i in enumerate(sel):
            assert nkeep[k] == g["nkeep"][i]
            assert list(keep[k, :nkeep[k]]) == list(g["keep"][i, :g["nkeep"][i]]), i


def test_keras_yolo3_adapter_vs_reference_fixture():
    """tools.yolo.YOLO mirror (k_yolo3_post) against the unmodified tools/yolo.py: boxes (transposed like the reference's),
    labels and scores bit for bit, batched and through detect_image with a stand-in model."""
    from PIL import Image
    from deepdish_b200.tools.yolo import YOLO
    g = goldens.load("yolo3.npz")
    net = int(g["net"])
    det = YOLO(wanted_labels=list(g["wanted"]), score_threshold=float(g["thr"]), model_image_size=(net, net),
               class_names=list(g["names"]))
    maps = [np.stack([g["map%d_%d" % (f, k)] for f in range(12)]) for k in range(3)]
    res = det.detect_maps(maps, (640, 480))
    total = 0
    for f, (boxes, labels, scores) in enumerate(res):
        np.testing.assert_array_equal(np.array(boxes, np.int64).reshape(-1, 4), g["box%d" % f])
        assert [list(g["names"]).index(l) for l in labels] == list(g["lab%d" % f])
        np.testing.assert_array_equal(np.array(scores, np.float32).view(np.uint32), g["score%d" % f].view(np.uint32))
        total += len(labels)
    assert total > 300

    class Model:
        def predict(self, data):
            assert data.shape == (1, net, net, 3) and data.dtype == np.float32
            return [g["map3_%d" % k][None] for k in range(3)]

    det.model = Model()
    boxes, labels, scores = det.detect_image(Image.new("RGB", (640, 480)))
    np.testing.assert_array_equal(np.array(boxes, np.int64).reshape(-1, 4), g["box3"])


def test_box_filter_vs_reference_fixture():
    """dd_box_filter against the reference's own loop (deepdish.py:941-960 inside Pipeline.detect_objects)."""
    from deepdish_b200 import ops
    g = goldens.load("box_filter.npz")
    out, idx, cnt = ops.box_filter(ops._dev(g["boxes"], torch.float64), ops._dev(g["counts"], torch.int32))
    out, idx, cnt = out.cpu().numpy(), idx.cpu().numpy(), cnt.cpu().numpy()
    for c in range(len(g["counts"])):
        k = len(g["kept%d" % c])
        assert cnt[c] == k, c
        np.testing.assert_array_equal(out[c, :k], g["out%d" % c].astype(float).reshape(-1, 4))
        np.testing.assert_array_equal(idx[c, :k], g["kept%d" % c])


def test_nms_tied_scores_vs_reference_fixture():
    """Tied scores (quantised heads produce them) at n <= 16, and scores=None (rank by y2): the reference's pick order
    -- higher index first among equals -- through the deep_sort.preprocessing mirror."""
    from deepdish_b200.deep_sort import preprocessing
    g = goldens.load("nms_ties.npz")
    ties = 0
    for i in range(len(g["counts"])):
        n = int(g["counts"][i])
        sc = g["scores"][i, :n] if g["use_scores"][i] else None
        keep = preprocessing.non_max_suppression(g["boxes"][i, :n].astype(np.int64), float(g["thr"][i]), sc)
        assert keep == list(g["keep"][i, :g["nkeep"][i]]), i
        ties += int(len(np.unique(g["scores"][i, :n])) < n)
    assert ties > 40


def test_yolo_decode_vs_reference_fixture():
    from deepdish_b200 import ops
    g = goldens.load("yolo.npz")
    names, wanted = list(g["names"]), list(g["wanted"])
    mask = torch.tensor([1 if n in wanted else 0 for n in names], dtype=torch.uint8, device="cuda")
    head = torch.from_numpy(g["head"]).cuda()
    out = ops.yolo_decode(head, mask, 0.25, (640, 480), (640, 480), ncap=256)
    keep, nkeep = ops.nms(out["tlwh"], out["score"], out["count"], 0.6)
    o = {k: v.cpu().numpy() for k, v in out.items()}
    keep, nkeep = keep.cpu().numpy(), nkeep.cpu().numpy()
    for f in range(head.shape[0]):
        n = int(o["count"][f])
        fidx = g["fidx%d" % f]                       # rows of the reference output that pass the box filter
        assert n == len(fidx)
        np.testing.assert_array_equal(o["tlwh"][f, :n], g["fbox%d" % f].reshape(-1, 4).astype(np.float64))
        np.testing.assert_array_equal(o["score"][f, :n], g["score%d" % f][fidx])
        np.testing.assert_array_equal(o["cls"][f, :n], g["cls%d" % f][fidx])
        assert np.all(np.diff(o["anchor"][f, :n]) > 0)
        assert list(keep[f, :nkeep[f]]) == list(g["keep%d" % f])
    assert int(o["flags"].sum()) == 0


def test_kalman_metric_iou_vs_reference_fixture():
    from deepdish_b200 import ops
    g = goldens.load("kalman.npz")
    mean, cov = ops.kalman_initiate(ops._dev(g["z0"], torch.float64))
    np.testing.assert_array_equal(mean.cpu().numpy(), g["seq_mean"][0])
    np.testing.assert_array_equal(cov.cpu().numpy(), g["seq_cov"][0])
    k = 0
    for it in range(5):
        ops.kalman_predict_(mean, cov); k += 1
        np.testing.assert_allclose(mean.cpu().numpy(), g["seq_mean"][k], rtol=1e-9)
        np.testing.assert_allclose(cov.cpu().numpy(), g["seq_cov"][k], rtol=1e-7, atol=1e-13)
        ops.kalman_update_(mean, cov, ops._dev(g["zs"][it], torch.float64)); k += 1
        np.testing.assert_allclose(mean.cpu().numpy(), g["seq_mean"][k], rtol=1e-9)
        np.testing.assert_allclose(cov.cpu().numpy(), g["seq_cov"][k], rtol=1e-7, atol=1e-13)
    meas = ops._dev(g["meas"], torch.float64)
    np.testing.assert_allclose(ops.kalman_gating_distance(mean, cov, meas).cpu().numpy(), g["gating4"], rtol=1e-7)
    np.testing.assert_allclose(ops.kalman_gating_distance(mean, cov, meas, True).cpu().numpy(), g["gating2"], rtol=1e-7)
    m = goldens.load("metric_iou.npz")
    gal, off, feats = ops._dev(m["gallery"], torch.float32), ops._dev(m["offsets"], torch.int32), ops._dev(m["feats"], torch.float32)
    np.testing.assert_allclose(ops.nn_distance(gal, off, feats, "cosine").cpu().numpy(), m["cosine"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(ops.nn_distance(gal, off, feats, "euclidean").cpu().numpy(), m["euclidean"], rtol=1e-4, atol=5e-4)
    cost = ops.iou_cost(ops._dev(m["trk_tlwh"], torch.float64), torch.ones(len(m["trk_tlwh"]), dtype=torch.int32, device="cuda"),
                        ops._dev(m["det_tlwh"], torch.float64))
    np.testing.assert_array_equal(cost.cpu().numpy(), 1. - m["iou"])


@pytest.mark.parametrize("n_chunks", [1, 3])
def test_batched_tracker_vs_multi_stream_reference_fixture(n_chunks):
    """Six cameras against six unmodified reference Trackers + Pipeline.process_results (tracker_multi.npz): the batched
    tracker -- stream chunks on their own CUDA streams, captured ticks through the native engine, alternating between
    HBM-resident batches and ragged pinned host batches -- reproduces every stream's det->track ids, track lists, states
    and counters tick by tick, and its count reduction equals the sum of the six reference counters."""
    from deepdish_b200.batched import BatchedTracker
    g = goldens.load("tracker_multi.npz")
    batches = goldens.multi_batches(g)
    if batches is None:
        pytest.skip("regenerated inputs do not match the fixture checksum")
    S, D = int(g["streams"]), int(g["dmax"])
    bt = BatchedTracker(S, LABELS3, max_tracks=96, max_dets=D, budget=int(g["budget"]), max_age=int(g["max_age"]),
                        n_chunks=n_chunks)
    ids_host = torch.empty((S, D), dtype=torch.int32).pin_memory()
    for f, b in enumerate(batches):
        if f % 2 == 0:
            bt.step(b.to("cuda"), join=False, reduce=True)
            bt.join()
            got = bt.det_track_id.cpu().numpy()
        else:
            bt.step_host_packed(bt.pack_host(b), ids_host)
            bt.join()
            torch.cuda.synchronize()
            got = ids_host.numpy().copy()
        np.testing.assert_array_equal(bt.total_counts.cpu().numpy(), g["counts"][:, f].sum(axis=0), err_msg="tick %d" % f)
        v = bt.host_view(["n_tracks", "order", "track_id", "state", "tsu", "n_deleted", "deleted", "counts"])
        for s in range(S):
            n = int(b.count[s])
            assert list(got[s, :n]) == list(g["det_ids"][s, f, :n]), (f, s)
            nt = int(v["n_tracks"][s]); sl = v["order"][s, :nt]
            assert nt == int(g["n_tracks"][s, f])
            assert list(v["track_id"][s, sl]) == list(g["ids"][s, f, :nt])
            assert list(v["state"][s, sl]) == list(g["states"][s, f, :nt])
            assert list(v["tsu"][s, sl]) == list(g["tsu"][s, f, :nt])
            dl = v["deleted"][s, :int(v["n_deleted"][s])]
            assert list(v["track_id"][s, dl]) == [x for x in g["deleted"][s, f] if x >= 0]
            np.testing.assert_array_equal(v["counts"][s], g["counts"][s, f])
    bt.check()
    vm = bt.host_view(["n_tracks", "order", "mean", "cov"])
    for s in range(S):
        nt = int(vm["n_tracks"][s]); sl = vm["order"][s, :nt]
        np.testing.assert_allclose(vm["mean"][s, sl], g["means"][s, :nt], rtol=1e-4, atol=1e-9)
        np.testing.assert_allclose(vm["cov"][s, sl], g["covs"][s, :nt], rtol=1e-4, atol=1e-12)


def test_yolo_decode_and_nms_full_size_vs_reference_fixture():
    """BASELINE configs[1] size: [frames, 25200, 85] head through dd_yolo_decode (TMA-staged tiles, fused box filter) and
    dd_nms against the unmodified reference's detect_image + box filter + NMS (yolo_full.npz): boxes, scores, classes
    and keep lists bit-exact."""
    from deepdish_b200 import ops
    g = goldens.load("yolo_full.npz")
    head_np = goldens.yolo_full_head(g)
    if head_np is None:
        pytest.skip("regenerated head does not match the fixture checksum")
    names, wanted = list(g["names"]), list(g["wanted"])
    mask = torch.tensor([1 if n in wanted else 0 for n in names], dtype=torch.uint8, device="cuda")
    head = torch.from_numpy(head_np).cuda()
    out = ops.yolo_decode(head, mask, 0.25, (640, 480), (640, 480), ncap=1024)
    keep, nkeep = ops.nms(out["tlwh"], out["score"], out["count"], 0.6)
    o = {k: v.cpu().numpy() for k, v in out.items()}
    keep, nkeep = keep.cpu().numpy(), nkeep.cpu().numpy()
    for f in range(head.shape[0]):
        n = int(o["count"][f])
        assert n == len(g["fbox%d" % f]) and n > 200
        np.testing.assert_array_equal(o["tlwh"][f, :n], g["fbox%d" % f])
        np.testing.assert_array_equal(o["score"][f, :n], g["fscore%d" % f])
        np.testing.assert_array_equal(o["cls"][f, :n], g["fcls%d" % f])
        assert list(keep[f, :nkeep[f]]) == list(g["keep%d" % f])
    assert int(o["flags"].sum()) == 0
