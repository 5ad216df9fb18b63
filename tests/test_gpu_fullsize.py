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
    _lib.check(_lib.lib().dd_tuning_set(0, 0), "dd_tuning_set")
    try:
        b3 = BatchedTracker(S, LABELS3, max_tracks=TMAX, max_dets=DMAX, budget=BUDGET, max_age=MAX_AGE, n_chunks=3)
        for f, b in enumerate(frames):
            b3.step(b, join=False, reduce=True)
            b3.join()
            np.testing.assert_array_equal(b3.det_track_id.cpu().numpy(), ids_a[f], err_msg="tick %d" % f)
        b3.check()
        vb = b3.host_view()
        np.testing.assert_array_equal(b3.total_counts.cpu().numpy(), tot_a)
    finally:
        _lib.check(_lib.lib().dd_tuning_set(0, 3), "dd_tuning_set")
    for name in ("n_tracks", "next_id", "n_deleted", "counts"):
        np.testing.assert_array_equal(va[name], vb[name], err_msg=name)
    live = np.arange(TMAX)[None, :] < va["n_tracks"][:, None]
    np.testing.assert_array_equal(va["order"][live], vb["order"][live])
    rows = np.repeat(np.arange(S), va["n_tracks"])
    slots = va["order"][live]
    for name in ("track_id", "state", "hits", "age", "tsu", "gal_len", "gal_pos", "mean", "cov"):
        np.testing.assert_array_equal(va[name][rows, slots], vb[name][rows, slots], err_msg=name)   # bit for bit
