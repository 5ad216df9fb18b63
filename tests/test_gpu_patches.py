"""GPU parity of the encoder-input step: extract_image_patch + DummyImageEncoder through the
C ABI, against the fixtures the unmodified reference produced and against the oracle at batch sizes."""
import numpy as np
import pytest
import torch

from oracle import patches as op
from tests import goldens

pytestmark = pytest.mark.gpu


def test_patches_golden_int_and_float_boxes():
    from deepdish_b200 import ops
    g = goldens.load("patches.npz")
    frame = torch.from_numpy(g["image"]).cuda()[None]
    for bk, pk, vk, is_int in (("boxes", "patches", "valid", True), ("fboxes", "fpatches", "fvalid", False)):
        boxes = torch.from_numpy(g[bk].astype(np.float64)).cuda()[None]
        patches, valid = ops.extract_patches(frame, boxes, None, (128, 64), boxes_are_int=is_int)
        np.testing.assert_array_equal(valid[0].cpu().numpy(), g[vk])
        np.testing.assert_array_equal(patches[0].cpu().numpy(), g[pk])


def test_dummy_encoder_golden_and_api_mirror():
    from deepdish_b200.tools import generate_detections as gd
    g = goldens.load("patches.npz")
    enc = gd.create_box_encoder("dummy")
    feat = enc(g["image"], list(g["dboxes"]))
    assert feat.dtype == np.float32
    np.testing.assert_array_equal(feat, g["dfeat"])                                   # bit-exact float32
    np.testing.assert_array_equal(gd.DummyImageEncoder()(g["flat"]), g["dflat"])       # zero row -> e0
    assert gd.extract_image_patch(g["image"], g["boxes"][1], (128, 64)) is None
    np.testing.assert_array_equal(gd.extract_image_patch(g["image"], g["boxes"][3], (128, 64)), g["patches"][3])
    assert enc(g["image"], []).size == 0
    c = gd.create_box_encoder("constant")(g["image"], list(g["dboxes"]))
    np.testing.assert_array_equal(c, op.constant_encode(len(g["dboxes"])))


@pytest.mark.parametrize("shape", [(128, 64), (16, 8), (32, 32), (7, 12)])
def test_patches_batched_vs_oracle(shape):
    """8 frames x 24 ragged boxes (incl. out-of-frame, 1-pixel, frame-sized), several patch shapes."""
    from deepdish_b200 import ops
    rng = np.random.default_rng(21)
    B, D, H, W = 8, 24, 480, 640
    frames = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    boxes = np.stack([rng.integers(-40, W + 20, (B, D)), rng.integers(-40, H + 20, (B, D)),
                      rng.integers(1, 200, (B, D)), rng.integers(1, 300, (B, D))], -1).astype(np.int64)
    boxes[0, 0] = [0, 0, W, H]
    boxes[0, 1] = [100, 100, 1, 1]
    counts = rng.integers(0, D + 1, B).astype(np.int32)
    counts[0] = D
    patches, valid = ops.extract_patches(torch.from_numpy(frames).cuda(), torch.from_numpy(boxes.astype(np.float64)).cuda(),
                                         torch.from_numpy(counts).cuda(), shape)
    patches, valid = patches.cpu().numpy(), valid.cpu().numpy()
    n_valid = 0
    for f in range(B):
        for d in range(D):
            if d >= counts[f]:
                assert valid[f, d] == 0
                continue
            exp = op.extract_image_patch(frames[f], boxes[f, d], shape)
            assert (exp is not None) == bool(valid[f, d]), (f, d)
            if exp is not None:
                n_valid += 1
                np.testing.assert_array_equal(patches[f, d], exp, err_msg="%d %d %s" % (f, d, boxes[f, d]))
            else:
                assert not patches[f, d].any()
    assert n_valid > 50


def test_patch_argument_errors():
    from deepdish_b200 import ops
    fr = torch.zeros((1, 8, 8, 3), dtype=torch.uint8, device="cuda")
    bx = torch.zeros((1, 2, 4), dtype=torch.float64, device="cuda")
    with pytest.raises(ValueError):
        ops.extract_patches(fr, bx, None, (16, 6))          # width not a multiple of 4
    with pytest.raises(ValueError):
        ops.dummy_encode(torch.zeros((2, 16, 8, 4), dtype=torch.uint8, device="cuda"))
