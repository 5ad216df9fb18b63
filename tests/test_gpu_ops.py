"""GPU parity of the stand-alone C-ABI operators against the oracle / scipy / CPython."""
import random

import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment

from oracle import deepsort as od, countline as oc

pytestmark = pytest.mark.gpu


_KEEP = []


def dev(a, dt):
    """Host array -> device tensor that stays alive (its data_ptr is handed to an async kernel)."""
    t = torch.as_tensor(np.ascontiguousarray(a), dtype=dt).cuda().contiguous()
    _KEEP.append(t)
    if len(_KEEP) > 256:
        torch.cuda.synchronize()
        del _KEEP[:128]
    return t


def L():
    from deepdish_b200 import _lib
    return _lib, _lib.lib()


def test_lsap_matches_scipy_on_tie_heavy_matrices():
    _lib, lib = L()
    rng = np.random.default_rng(1)
    for (nr, nc) in [(1, 1), (3, 7), (7, 3), (12, 12), (50, 50), (64, 37), (40, 200), (256, 200)]:
        B = 48 if nr * nc < 10000 else 6
        cost = np.empty((B, nr, nc))
        for b in range(B):
            k = b % 4
            if k == 0:
                c = rng.integers(0, 4, (nr, nc)).astype(float)
            elif k == 1:
                c = rng.random((nr, nc)); c[c > 0.3] = 0.2 + 1e-5
            elif k == 2:
                c = np.full((nr, nc), 0.2 + 1e-5)
                for i in range(min(nr, nc)):
                    if rng.random() < 0.7:
                        c[i, rng.integers(0, nc)] = rng.random() * 0.2
            else:
                c = rng.random((nr, nc))
            cost[b] = c
        out = torch.full((B, nr), -7, dtype=torch.int32, device="cuda")
        st = torch.zeros(B, dtype=torch.int32, device="cuda")
        _lib.check(lib.dd_lsap(dev(cost, torch.float64).data_ptr(), B, nr, nc, out.data_ptr(), st.data_ptr(), None), "lsap")
        got = out.cpu().numpy()
        assert int(st.sum()) == 0
        for b in range(B):
            r, c = linear_sum_assignment(cost[b])
            exp = np.full(nr, -1); exp[r] = c
            np.testing.assert_array_equal(got[b], exp, err_msg=str((nr, nc, b)))


def test_set_difference_order_matches_cpython():
    _lib, lib = L()
    random.seed(3)
    B, NA = 4000, 256
    a = np.zeros((B, NA), np.int32); na = np.zeros(B, np.int32)
    m = np.zeros((B, NA), np.int32); nm = np.zeros(B, np.int32)
    exp = []
    for b in range(B):
        n = random.randint(0, NA)
        av = list(range(n)) if b % 5 else random.sample(range(1000), n)
        k = random.randint(0, n) if b % 3 else min(n, random.randint(0, 6))
        mv = random.sample(av, k)
        a[b, :n] = av; na[b] = n; m[b, :k] = mv; nm[b] = k
        exp.append(list(set(av) - set(mv)))
    out = torch.zeros((B, NA), dtype=torch.int32, device="cuda")
    on = torch.zeros(B, dtype=torch.int32, device="cuda")
    _lib.check(lib.dd_set_difference_order(dev(a, torch.int32).data_ptr(), dev(na, torch.int32).data_ptr(), NA,
                                           dev(m, torch.int32).data_ptr(), dev(nm, torch.int32).data_ptr(), NA, B,
                                           out.data_ptr(), on.data_ptr(), None), "setdiff")
    o, n = out.cpu().numpy(), on.cpu().numpy()
    for b in range(B):
        assert list(o[b, :n[b]]) == exp[b], b


def test_kalman_ops():
    _lib, lib = L()
    rng = np.random.default_rng(0)
    n, m = 300, 70
    xyah = np.c_[rng.uniform(0, 600, n), rng.uniform(0, 400, n), rng.uniform(0.2, 0.8, n), rng.uniform(40, 100, n)]
    mean = torch.zeros((n, 8), dtype=torch.float64, device="cuda")
    cov = torch.zeros((n, 8, 8), dtype=torch.float64, device="cuda")
    _lib.check(lib.dd_kalman_initiate(dev(xyah, torch.float64).data_ptr(), mean.data_ptr(), cov.data_ptr(), n, None), "init")
    om, ocv = zip(*[od.kf_initiate(z) for z in xyah])
    np.testing.assert_array_equal(mean.cpu().numpy(), np.stack(om))
    np.testing.assert_array_equal(cov.cpu().numpy(), np.stack(ocv))
    om, ocv = list(om), list(ocv)
    for it in range(6):
        _lib.check(lib.dd_kalman_predict(mean.data_ptr(), cov.data_ptr(), n, None), "predict")
        for i in range(n):
            om[i], ocv[i] = od.kf_predict(om[i], ocv[i])
        np.testing.assert_allclose(mean.cpu().numpy(), np.stack(om), rtol=1e-12)
        np.testing.assert_allclose(cov.cpu().numpy(), np.stack(ocv), rtol=1e-10, atol=1e-14)
        z = np.stack([om[i][:4] + rng.normal(0, 1, 4) * [2, 2, 0.01, 2] for i in range(n)])
        if it % 2 == 0:
            _lib.check(lib.dd_kalman_update(mean.data_ptr(), cov.data_ptr(), dev(z, torch.float64).data_ptr(), n, None), "update")
            for i in range(n):
                om[i], ocv[i] = od.kf_update(om[i], ocv[i], z[i])
            np.testing.assert_allclose(mean.cpu().numpy(), np.stack(om), rtol=1e-9)
            np.testing.assert_allclose(cov.cpu().numpy(), np.stack(ocv), rtol=1e-7, atol=1e-12)
    pm = torch.zeros((n, 4), dtype=torch.float64, device="cuda")
    pc = torch.zeros((n, 4, 4), dtype=torch.float64, device="cuda")
    _lib.check(lib.dd_kalman_project(mean.data_ptr(), cov.data_ptr(), pm.data_ptr(), pc.data_ptr(), n, None), "project")
    mh, ch = mean.cpu().numpy(), cov.cpu().numpy()
    for i in range(0, n, 17):
        a, b = od.kf_project(mh[i], ch[i])
        np.testing.assert_array_equal(pm[i].cpu().numpy(), a)
        np.testing.assert_array_equal(pc[i].cpu().numpy(), b)
    meas = np.c_[rng.uniform(0, 600, m), rng.uniform(0, 400, m), rng.uniform(0.2, 0.8, m), rng.uniform(40, 100, m)]
    meas[:n // 8] = mh[:n // 8 * 1, :4][: len(meas[:n // 8])] + 0.5
    for only_pos in (0, 1):
        out = torch.zeros((n, m), dtype=torch.float64, device="cuda")
        _lib.check(lib.dd_kalman_gating_distance(mean.data_ptr(), cov.data_ptr(), dev(meas, torch.float64).data_ptr(),
                                                 n, m, only_pos, out.data_ptr(), None), "gating")
        exp = np.stack([od.kf_gating_distance(mh[i], ch[i], meas, bool(only_pos)) for i in range(n)])
        np.testing.assert_allclose(out.cpu().numpy(), exp, rtol=1e-8)


def test_nn_distance_and_iou():
    _lib, lib = L()
    rng = np.random.default_rng(2)
    lens = [1, 3, 100, 17, 64]
    gal = [rng.normal(size=(l, 128)).astype(np.float32) * rng.uniform(0.5, 2) for l in lens]
    feats = rng.normal(size=(23, 128)).astype(np.float32)
    feats[3] = gal[2][40] + 0.01 * rng.normal(size=128).astype(np.float32)
    off = np.r_[0, np.cumsum(lens)].astype(np.int32)
    for metric, fn in ((0, od.nn_cosine_distance), (1, od.nn_euclidean_distance)):
        out = torch.zeros((len(lens), 23), dtype=torch.float64, device="cuda")
        _lib.check(lib.dd_nn_distance(dev(np.concatenate(gal), torch.float32).data_ptr(), dev(off, torch.int32).data_ptr(),
                                      dev(feats, torch.float32).data_ptr(), len(lens), 23, metric, out.data_ptr(), None), "nn")
        exp = np.stack([fn(g, feats) for g in gal]).astype(np.float64)
        # cosine: 1e-4 relative + the reference's own f32 summation noise (~1e-7 abs on a unit dot);
        # euclidean: |a|^2 + |b|^2 - 2ab cancels ~5e2-magnitude f32 terms -> noise ~1e-7 * 5e2 abs
        atol = 2e-6 if metric == 0 else 5e-4
        np.testing.assert_allclose(out.cpu().numpy(), exp, rtol=1e-4, atol=atol)
    n, m = 40, 33
    trk = np.c_[rng.uniform(0, 500, n), rng.uniform(0, 300, n), rng.uniform(20, 80, n), rng.uniform(40, 120, n)]
    det = np.floor(np.c_[rng.uniform(0, 500, m), rng.uniform(0, 300, m), rng.uniform(20, 80, m), rng.uniform(40, 120, m)])
    det[:10] = np.floor(trk[:10]) + 1
    tsu = (rng.random(n) < 0.2).astype(np.int32) + 1
    out = torch.zeros((n, m), dtype=torch.float64, device="cuda")
    _lib.check(lib.dd_iou_cost(dev(trk, torch.float64).data_ptr(), dev(tsu, torch.int32).data_ptr(),
                               dev(det, torch.float64).data_ptr(), n, m, out.data_ptr(), None), "iou")
    exp = np.stack([np.full(m, od.INFTY_COST) if tsu[i] > 1 else 1. - od.iou(trk[i], det) for i in range(n)])
    np.testing.assert_array_equal(out.cpu().numpy(), exp)


def test_intersection_golden_and_random():
    _lib, lib = L()
    from tests.test_oracle_intersection import GOLDEN_SEGMENTS
    segs = [np.r_[p, pr, q, qs] for (p, pr, q, qs, _) in GOLDEN_SEGMENTS]
    exp = [e for (*_, e) in GOLDEN_SEGMENTS]
    rng = np.random.default_rng(4)
    for _ in range(5000):
        s = rng.integers(0, 6, 8).astype(float) if rng.random() < 0.5 else rng.uniform(0, 5, 8)
        segs.append(s)
        exp.append(oc.intersection(s[0:2], s[2:4], s[4:6], s[6:8]))
    out = torch.zeros(len(segs), dtype=torch.int32, device="cuda")
    _lib.check(lib.dd_intersection(dev(np.stack(segs), torch.float64).data_ptr(), len(segs), out.data_ptr(), None), "isect")
    np.testing.assert_array_equal(out.cpu().numpy().astype(bool), np.array(exp, dtype=bool))
