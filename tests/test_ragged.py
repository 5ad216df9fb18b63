"""Host side of the ragged detection batch (deepdish_b200/ragged.py): layout, alignment and round trip."""
import numpy as np
import pytest

from deepdish_b200 import ragged


def _batch(rng, n, D, counts):
    tlwh = rng.integers(0, 600, (n, D, 4)).astype(np.float64)
    conf = rng.uniform(0.5, 1, (n, D)).astype(np.float32)
    label = rng.integers(0, 3, (n, D)).astype(np.int32)
    feat = rng.normal(size=(n, D, 128)).astype(np.float32)
    return tlwh, conf, label, feat, np.asarray(counts, np.int32)


@pytest.mark.parametrize("counts", [[3, 0, 5, 1], [0, 0, 0], [6, 6], [1]])
def test_pack_round_trip_and_alignment(counts):
    rng = np.random.default_rng(1)
    n, D = len(counts), 6
    tlwh, conf, label, feat, cnt = _batch(rng, n, D, counts)
    blob, total, (o_tlwh, o_conf, o_label, o_feat) = ragged.pack(tlwh, conf, label, feat, cnt)
    N = int(sum(counts))
    assert total == o_feat + 512 * N and blob.size >= total
    assert o_tlwh % 16 == 0 and o_feat % 16 == 0 and o_conf % 4 == 0 and o_label % 4 == 0       # what the C ABI requires
    assert list(blob[:4 * (n + 1)].view(np.int32)) == [0] + list(np.cumsum(counts))
    t2, c2, l2, f2, k2 = ragged.unpack(blob, n, D)
    np.testing.assert_array_equal(k2, cnt)
    sel = np.arange(D)[None, :] < cnt[:, None]
    for got, exp in ((t2, tlwh), (c2, conf), (l2, label), (f2, feat)):
        np.testing.assert_array_equal(got[sel], exp[sel])
        assert not got[~sel].any()
    # padding is never uploaded: the blob is exactly as large as the detections that exist
    assert total <= 16 + 4 * (n + 1) + 16 + N * (32 + 4 + 4 + 512) + 16


def test_pack_rejects_bad_counts_and_small_buffers():
    rng = np.random.default_rng(2)
    tlwh, conf, label, feat, cnt = _batch(rng, 2, 4, [5, 1])
    with pytest.raises(ValueError):
        ragged.pack(tlwh, conf, label, feat, cnt)
    cnt[0] = 4
    with pytest.raises(ValueError):
        ragged.pack(tlwh, conf, label, feat, cnt, out=np.zeros(64, np.uint8))
