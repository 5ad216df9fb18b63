"""CPU restatement of the patch extraction that feeds the re-ID encoder.
TEST INFRASTRUCTURE.

Follows tools/generate_detections.py:40-84 (extract_image_patch) and :198-205 (the encoder's loop over
boxes), paths relative to /root/reference.  The resize is OpenCV's 8-bit bilinear ``cv2.resize`` (third-party;
the installed version here is 4.13.0): restated from its fixed-point scheme (11-bit coefficients, horizontal
pass in int32, vertical pass ``((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2``) and checked bit-exact
against cv2 itself in tests/test_golden.py.  Pinned: tests/golden/patches.npz holds patches the unmodified
reference function produced.
"""
import numpy as np


def patch_box(bbox, patch_shape, img_h, img_w):
    """generate_detections.py:64-80 for an INTEGER box array (deepdish.py:993 hands int64 boxes over, so the
    in-place float updates of :69-70 truncate toward zero).  Returns (sx, sy, ex, ey) or None."""
    x, y, w, h = (int(v) for v in bbox)
    target_aspect = float(patch_shape[1]) / patch_shape[0]
    new_width = target_aspect * h
    x = int(x - (new_width - w) / 2)            # bbox[0] -= ... on an int64 array: C-style truncation
    w = int(new_width)
    ex, ey = x + w, y + h
    sx, sy = max(0, x), max(0, y)
    ex, ey = min(img_w - 1, ex), min(img_h - 1, ey)
    if sx >= ex or sy >= ey:
        return None
    return sx, sy, ex, ey


def patch_box_float(bbox, patch_shape, img_h, img_w):
    """generate_detections.py:64-80 for a FLOAT box array: nothing truncates until ``astype(np.int)``
    (:73), which truncates x, y, x + new_width and y + h toward zero one by one."""
    x, y, w, h = (float(v) for v in bbox)
    target_aspect = float(patch_shape[1]) / patch_shape[0]
    new_width = target_aspect * h
    x = x - (new_width - w) / 2
    ex, ey = int(new_width + x), int(h + y)        # bbox[2:] += bbox[:2]
    sx, sy = max(0, int(x)), max(0, int(y))
    ex, ey = min(img_w - 1, ex), min(img_h - 1, ey)
    if sx >= ex or sy >= ey:
        return None
    return sx, sy, ex, ey


def _coeffs(dn, sn, clamp):
    scale = 1.0 / (dn / sn)
    i0 = np.zeros(dn, np.int64); i1 = np.zeros(dn, np.int64)
    a0 = np.zeros(dn, np.int64); a1 = np.zeros(dn, np.int64)
    for d in range(dn):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - s)
        if clamp:                                # x: cv2 zeroes the fraction at the borders
            if s < 0:
                s, f = 0, np.float32(0)
            if s >= sn - 1:
                s, f = sn - 1, np.float32(0)
            i0[d], i1[d] = s, min(s + 1, sn - 1)
        else:                                    # y: cv2 clips the row indices, keeps the weights
            i0[d], i1[d] = min(max(s, 0), sn - 1), min(max(s + 1, 0), sn - 1)
        a0[d] = int(np.rint(np.float32((np.float32(1) - f) * np.float32(2048))))
        a1[d] = int(np.rint(np.float32(f * np.float32(2048))))
    return i0, i1, a0, a1


def resize_bilinear_u8(src, dw, dh):
    """cv2.resize(src, (dw, dh)) for uint8 images, INTER_LINEAR, restated."""
    sh, sw = src.shape[:2]
    xi, xi1, xa0, xa1 = _coeffs(dw, sw, True)
    yi, yi1, ya0, ya1 = _coeffs(dh, sh, False)
    s = src.astype(np.int64)
    rows = s[:, xi, :] * xa0[None, :, None] + s[:, xi1, :] * xa1[None, :, None]
    s0, s1 = rows[yi], rows[yi1]
    out = (((ya0[:, None, None] * (s0 >> 4)) >> 16) + ((ya1[:, None, None] * (s1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def extract_image_patch(image, bbox, patch_shape):
    """generate_detections.py:40-84 -> uint8 patch [ph, pw, c] or None."""
    fn = patch_box if np.issubdtype(np.asarray(bbox).dtype, np.integer) else patch_box_float
    box = fn(bbox, patch_shape, image.shape[0], image.shape[1])
    if box is None:
        return None
    sx, sy, ex, ey = box
    return resize_bilinear_u8(image[sy:ey, sx:ex], patch_shape[1], patch_shape[0])


def dummy_encode(patches):
    """DummyImageEncoder.__call__ (generate_detections.py:92-105): patches u8 [n,16,8,3] -> f32 [n,128].
    Restated with explicit float32 steps: channel mean (sums of three integers <= 765 are exact, one
    correctly rounded division by 3), minus 128, L2 norm with numpy's pairwise float32 summation of 128
    squares (eight strided accumulators, combined ((0+1)+(2+3))+((4+5)+(6+7))), one rounded sqrt and one
    rounded division per element; an all-zero row becomes e0."""
    p = np.asarray(patches)
    n = p.shape[0]
    f = np.float32
    s = (p[..., 0].astype(f) + p[..., 1].astype(f)) + p[..., 2].astype(f)
    mat = (s / f(3)).astype(f).reshape(n, 128) - f(128)
    out = np.zeros((n, 128), f)
    for i in range(n):
        sq = (mat[i] * mat[i]).astype(f)
        r = [f(0)] * 8
        for j in range(8):
            acc = sq[j]
            for k in range(1, 16):
                acc = f(acc + sq[8 * k + j])
            r[j] = acc
        tot = f(f(f(r[0] + r[1]) + f(r[2] + r[3])) + f(f(r[4] + r[5]) + f(r[6] + r[7])))
        l = f(np.sqrt(tot))
        if l == 0:
            out[i] = mat[i]
            out[i, 0] = 1
        else:
            out[i] = mat[i] / l
    return out


def constant_encode(n):
    """ConstantImageEncoder.__call__ (generate_detections.py:113-116)."""
    out = np.zeros((n, 128), np.float32)
    out[:, 0] = 1
    return out
