"""CPU tests of the kernel bodies' serial logic through the host emulation (tests/hostemu): the same
source files the CUDA kernels are built from, compiled with a one-lane group.  Not a product path."""
import random

import numpy as np
import pytest
from scipy.optimize import linear_sum_assignment

from deepdish_b200 import _lib as L
from deepdish_b200.scene import Scene
from oracle import deepsort as od
from tests.hostemu import driver as hd
from tests.parity import OracleStreams, compare_stream, compare_costs, LABELS3


def _matrices(rng, n):
    for trial in range(n):
        nr, nc = rng.integers(1, 14), rng.integers(1, 14)
        kind = trial % 4
        if kind == 0:
            c = rng.integers(0, 4, (nr, nc)).astype(float)
        elif kind == 1:
            c = rng.random((nr, nc)); c[c > 0.3] = 0.20001
        elif kind == 2:
            c = np.full((nr, nc), 0.20001)
            for i in range(min(nr, nc)):
                if rng.random() < 0.7:
                    c[i, rng.integers(0, nc)] = rng.random() * 0.2
        else:
            c = rng.random((nr, nc))
        yield c


def test_lsap_device_code_and_oracle_port_match_scipy():
    rng = np.random.default_rng(0)
    for c in _matrices(rng, 1500):
        r, cc = linear_sum_assignment(c)
        exp = np.full(c.shape[0], -1); exp[r] = cc
        rc, out = hd.lsap(c)
        assert rc == 0 and np.array_equal(exp, out)
        pr, pc = od.lsap_port(c)
        assert np.array_equal(pr, r) and np.array_equal(pc, cc)


def test_set_difference_order_matches_cpython():
    random.seed(1)
    for trial in range(6000):
        n = random.randint(0, 200)
        a = list(range(n)) if trial % 5 else random.sample(range(300), min(n, 250))
        m = random.sample(a, random.randint(0, len(a))) if trial % 3 else random.sample(a, min(len(a), random.randint(0, 6)))
        exp = list(set(a) - set(m))
        assert hd.set_difference_order(a, m) == exp
        if trial % 5:
            assert od.cpython_set_difference_order(n, m) == exp


def test_trajectory_parity_host_emulation():
    S = 3
    cfg = L.make_config(S, 64, 24, 100, LABELS3, max_age=30)
    emu = hd.HostEmuTracker(cfg)
    orc = OracleStreams(S, LABELS3, max_age=30)
    sc = Scene(S, 20, 24, n_labels=3, seed=7)
    checked = 0
    for f in range(70):
        b = sc.step()
        pre = {k: emu.v[k].copy() for k in ("n_tracks", "order", "track_id")}
        ids = orc.step(b)
        emu.predict()
        got = emu.update(b.tlwh.numpy(), b.conf.numpy(), b.label.numpy(), b.feat.numpy(), b.count.numpy())
        emu.countline()
        for s in range(S):
            n = int(b.count[s])
            assert list(got[s, :n]) == ids[s], (f, s)
            compare_stream(orc.trk[s], orc.cnt[s], emu.v, s, LABELS3, gallery_rows=emu.gallery if f % 10 == 9 else None)
            if f > 5:
                checked += compare_costs(orc.trk[s], pre, emu.v, s)
    assert checked > 500 and int(emu.v["err"].sum()) == 0
    # every page is either on the free stack or in exactly one live slot's table
    ctl, np_ = emu.v["pool_ctl"], emu.v["gal_np"]
    assert int(ctl[0]) + int(np_.sum()) == int(ctl[1])


def test_two_gate_words_host_emulation():
    """More than 32 detections per stream (two gate words per track) in a dense scene: the matching warp's cost staging
    walks a cursor across the words, tracks with more than DD_CVAL gate-passing detections read the rest of their costs
    from the global cost rows, and the unmatched-track set is read back from whichever table the growth left it in."""
    S = 2
    cfg = L.make_config(S, 96, 40, 30, LABELS3, max_age=12)
    emu = hd.HostEmuTracker(cfg)
    orc = OracleStreams(S, LABELS3, budget=30, max_age=12)
    sc = Scene(S, 34, 40, n_labels=3, seed=23)
    many, two_words, checked = 0, 0, 0
    for f in range(45):
        b = sc.step()
        pre = {k: emu.v[k].copy() for k in ("n_tracks", "order", "track_id")}
        ids = orc.step(b)
        emu.predict()
        got = emu.update(b.tlwh.numpy(), b.conf.numpy(), b.label.numpy(), b.feat.numpy(), b.count.numpy())
        emu.countline()
        gate = emu.v["gate"].view(np.uint32)
        pc = np.array([[bin(int(w)).count("1") for w in row] for row in gate.reshape(-1, gate.shape[-1])])
        many += int((pc.sum(axis=1) > 4).sum())
        two_words += int((pc[:, 1] > 0).sum())
        for s in range(S):
            n = int(b.count[s])
            assert list(got[s, :n]) == ids[s], (f, s)
            compare_stream(orc.trk[s], orc.cnt[s], emu.v, s, LABELS3)
            if f > 5:
                checked += compare_costs(orc.trk[s], pre, emu.v, s)
    assert int(emu.v["err"].sum()) == 0 and checked > 500
    assert many > 0 and two_words > 0, (many, two_words)       # the paths this test is for were taken


@pytest.mark.parametrize("budget", [None, 5, 16, 37])
def test_paged_galleries_host_emulation(budget):
    """nn_budget=None (deepdish.py:515-516: galleries never trimmed, nn_matching.py:137-154) and ring budgets that
    are smaller than / equal to / not a multiple of the 16-row page: ids, states, gallery contents and the page
    accounting against the oracle."""
    S = 2
    cfg = L.make_config(S, 48, 16, budget, LABELS3, max_age=8, page_cap=8)
    emu = hd.HostEmuTracker(cfg)
    orc = OracleStreams(S, LABELS3, budget=budget, max_age=8)
    sc = Scene(S, 10, 16, n_labels=3, seed=11)
    for f in range(100):
        b = sc.step()
        ids = orc.step(b)
        emu.predict()
        got = emu.update(b.tlwh.numpy(), b.conf.numpy(), b.label.numpy(), b.feat.numpy(), b.count.numpy())
        emu.countline()
        for s in range(S):
            n = int(b.count[s])
            assert list(got[s, :n]) == ids[s], (f, s)
            compare_stream(orc.trk[s], orc.cnt[s], emu.v, s, LABELS3, gallery_rows=emu.gallery if f % 7 == 6 else None)
    assert int(emu.v["err"].sum()) == 0
    ctl = emu.v["pool_ctl"]
    assert int(ctl[0]) + int(emu.v["gal_np"].sum()) == int(ctl[1])
    if budget is None:
        assert int(ctl[2]) > 64          # longest gallery so far (tracked for unbounded galleries only)
    else:
        assert int(emu.v["gal_len"].max()) == budget


def test_detection_bodies_vs_reference_fixtures():
    """NMS, YOLO row decode + box filter and the SSD post-processing device functions (host emulation)
    against the fixtures produced by the unmodified reference."""
    from tests import goldens
    g = goldens.load("nms.npz")
    for i in range(len(g["counts"])):
        n = int(g["counts"][i])
        assert hd.nms(g["boxes"][i, :n], g["scores"][i, :n], float(g["thr"][i])) == list(g["keep"][i, :g["nkeep"][i]])
    g = goldens.load("box_filter.npz")         # the reference's own filter loop (deepdish.py:941-960) on adversarial boxes
    for c in range(len(g["counts"])):
        out, idx = hd.box_filter(g["boxes"][c, :int(g["counts"][c])])
        np.testing.assert_array_equal(out, g["out%d" % c].astype(float).reshape(-1, 4), err_msg=str(c))
        np.testing.assert_array_equal(idx, g["kept%d" % c])
    g = goldens.load("yolo3.npz")              # Keras YOLOv3 adapter: unmodified tools/yolo.py (scalar-exp run), bit for bit
    names, wanted = list(g["names"]), list(g["wanted"])
    wm = np.array([1 if n in wanted else 0 for n in names], np.uint8)
    for f in range(12):
        b, s_, l, fl = hd.yolo3([g["map%d_%d" % (f, k)] for k in range(3)], g["anchors"], len(names), wm, float(g["thr"]),
                                (640, 480), (int(g["net"]), int(g["net"])))
        assert fl == 0
        np.testing.assert_array_equal(b, g["box%d" % f].astype(float))
        np.testing.assert_array_equal(l, g["lab%d" % f])
        np.testing.assert_array_equal(s_.view(np.uint32), g["score%d" % f].view(np.uint32))
    g = goldens.load("nms_ties.npz")           # tied scores: the reference picks the higher index first (n <= 16)
    for i in range(len(g["counts"])):
        n = int(g["counts"][i])
        b = g["boxes"][i, :n]
        sc = g["scores"][i, :n] if g["use_scores"][i] else (b[:, 1] + b[:, 3]).astype(np.float32)
        assert hd.nms(b, sc, float(g["thr"][i])) == list(g["keep"][i, :g["nkeep"][i]]), i
    g = goldens.load("yolo.npz")
    names, wanted = list(g["names"]), list(g["wanted"])
    mask = np.array([1 if n in wanted else 0 for n in names], np.uint8)
    for f in range(g["head"].shape[0]):
        tl, sc, cl, an, nan = hd.yolo_rows(g["head"][f], mask, 0.25, (640, 480), (640, 480))
        fidx = g["fidx%d" % f]
        assert nan == 0 and len(tl) == len(fidx)
        np.testing.assert_array_equal(tl, g["fbox%d" % f].reshape(-1, 4).astype(float))
        np.testing.assert_array_equal(sc, g["score%d" % f][fidx])
        np.testing.assert_array_equal(cl, g["cls%d" % f][fidx])
    g = goldens.load("ssd_post.npz")
    names, wanted = list(g["names"]), list(g["wanted"])
    c2l = np.array([(c + 1) if (c + 1 < len(names) and names[c + 1] in wanted) else -1 for c in range(len(names) - 1)], np.int32)
    total = 0
    for c in range(len(g["op_boxes"])):
        tl, sc, lb = hd.ssd_post(g["op_boxes"][c], g["op_classes"][c].astype(np.int32), g["op_scores"][c], c2l)
        # the fixture holds detect_image's float boxes and what the reference's own box filter makes of them; the
        # device function applies that filter too
        ib, kept = g["fbox%d" % c], g["fidx%d" % c]
        np.testing.assert_array_equal(tl, ib.astype(float).reshape(-1, 4))
        np.testing.assert_array_equal(sc, g["score%d" % c][kept])
        np.testing.assert_array_equal(lb, g["lab%d" % c][kept])
        total += len(tl)
    assert total > 100
