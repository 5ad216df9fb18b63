"""The gallery kernel's half-precision pre-pass must be invisible: its cost entries (min cosine distance of every
gate-passing (track, detection) pair) are compared BIT FOR BIT with the exact f32 pass on the same tracker history,
on scenes built to stress the re-check: identical gallery rows (every row inside the window), near-duplicates,
short galleries (fewer rows than one 16-row step), budgets that are not multiples of 16, more than 8 gate-passing
detections per track (several query groups) and more than 32 detections (several gate words)."""
import numpy as np
import pytest

from deepdish_b200.scene import Scene
from tests.parity import LABELS3

pytestmark = pytest.mark.gpu


def _history(impl, S, nobj, dmax, tmax, budget, frames, seed, feat_noise, big_boxes=False):
    from deepdish_b200.batched import BatchedTracker
    bt = BatchedTracker(S, LABELS3, max_tracks=tmax, max_dets=dmax, budget=budget, max_age=30, gallery_impl=impl,
                        page_cap=8)
    sc = Scene(S, nobj, dmax, n_labels=3, seed=seed, feat_noise=feat_noise)
    if big_boxes:                      # large slow boxes: many detections pass each track's Mahalanobis gate
        sc.size = sc.size * 4.0
        sc.vel = sc.vel * 0.2
    out = []
    for f in range(frames):
        b = sc.step().to("cuda")
        ids = bt.step(b).cpu().numpy().copy()
        bt.maintain(wait=True)
        gate = bt.v["gate"].cpu().numpy().astype(np.uint32)
        cost = bt.v["cost"].cpu().numpy()
        state = bt.v["state"].cpu().numpy()
        out.append((ids, gate, cost, state))
    bt.check()
    return out


@pytest.mark.parametrize("name,kw", [
    ("plain", dict(S=6, nobj=20, dmax=24, tmax=64, budget=100, frames=70, seed=41, feat_noise=0.02)),
    ("identical_rows", dict(S=4, nobj=12, dmax=16, tmax=48, budget=37, frames=60, seed=42, feat_noise=0.0)),
    ("near_duplicates", dict(S=4, nobj=12, dmax=16, tmax=48, budget=20, frames=50, seed=43, feat_noise=2e-4)),
    ("short_gallery_budget_5", dict(S=4, nobj=12, dmax=16, tmax=48, budget=5, frames=40, seed=44, feat_noise=0.02)),
    ("many_candidates_two_gate_words", dict(S=3, nobj=40, dmax=48, tmax=128, budget=33, frames=40, seed=45,
                                            feat_noise=0.05, big_boxes=True)),
    # identical rows x up to 8 gate-passing detections: more window candidates than a checker message lists (240), so
    # the checker's evaluate-everything fallback runs
    ("identical_rows_many_candidates", dict(S=3, nobj=30, dmax=40, tmax=96, budget=100, frames=115, seed=48,
                                            feat_noise=0.0, big_boxes=True)),
    # a crowd: more than 16 gate-passing detections per track (three and more jobs per gallery), five gate words, the
    # four-warp matching kernel
    ("crowd_many_jobs_per_track", dict(S=2, nobj=70, dmax=144, tmax=192, budget=33, frames=30, seed=49,
                                      feat_noise=0.05, big_boxes=True)),
    # nn_budget=None: galleries of ~300 rows span several 256-row blocks of the half pre-pass (running maximum)
    ("unbounded_multi_block", dict(S=2, nobj=8, dmax=12, tmax=32, budget=None, frames=340, seed=46, feat_noise=0.01)),
    ("unbounded_identical_rows", dict(S=2, nobj=6, dmax=8, tmax=32, budget=None, frames=300, seed=47, feat_noise=0.0)),
])
@pytest.mark.parametrize("impl", ["default"])
def test_half_prepass_costs_are_bit_identical_to_the_exact_pass(name, kw, impl):
    half = _history(impl, **kw)
    exact = _history("exact", **kw)
    checked = many = 0
    for f, ((ih, gh, ch, sh), (ie, ge, ce, se)) in enumerate(zip(half, exact)):
        np.testing.assert_array_equal(ih, ie, err_msg="%s ids frame %d" % (name, f))
        np.testing.assert_array_equal(gh, ge)
        S, T, D = ch.shape
        bits = ((gh[:, :, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(S, T, -1)[:, :, :D].astype(bool)
        # gate words of slots that were not confirmed at gating time are stale; ids equal => same tracks were gated.
        # Compare every entry both runs wrote this tick: identical bit patterns.
        a, b = ch.view(np.uint32), ce.view(np.uint32)
        np.testing.assert_array_equal(a[bits], b[bits], err_msg="%s cost bits frame %d" % (name, f))
        checked += int(bits.sum())
        many = max(many, int(bits.sum(axis=2).max()))
    assert checked > 500
    if name == "many_candidates_two_gate_words":
        assert many > 8
    if name == "identical_rows_many_candidates":
        assert many >= 3
    if name == "crowd_many_jobs_per_track":
        assert many > 16
