"""``tools.generate_detections`` mirror (reference tools/generate_detections.py): the box encoder that turns
(frame, boxes) into the 128-d appearance features ``Detection`` carries.

Built here: ``extract_image_patch`` (:40-84, CUDA kernel ``k_extract_patches``, bit-exact with the cv2.resize
underneath), the reference's two arithmetic encoders ``DummyImageEncoder`` (:86-105, kernel ``k_dummy_encode``)
and ``ConstantImageEncoder`` (:107-116), and ``create_box_encoder`` (:180-215).  The MARS CNN encoders
(``ImageEncoder`` / ``TFLiteImageEncoder``) are out of scope: pass ``image_encoder=`` a callable
``cuda uint8 patches [n,h,w,3] -> [n,feature_dim]`` with an ``image_shape`` attribute to plug one in.
"""
import numpy as np
import torch

from .. import ops


def _boxes_tensor(boxes):
    arr = np.asarray(boxes)
    is_int = np.issubdtype(arr.dtype, np.integer)
    return torch.as_tensor(arr.astype(np.float64)).reshape(1, -1, 4).cuda(), is_int


def extract_image_patch(image, bbox, patch_shape):
    """generate_detections.py:40-84 -> uint8 ndarray [h,w,3] or None (empty / fully outside box)."""
    if patch_shape is None:
        raise NotImplementedError("patch_shape=None (variable-size crops) is not used on the deepdish path")
    frame = torch.as_tensor(np.ascontiguousarray(image, dtype=np.uint8)).cuda()[None]
    boxes, is_int = _boxes_tensor(bbox)
    patches, valid = ops.extract_patches(frame, boxes, None, tuple(patch_shape[:2]), boxes_are_int=is_int)
    if int(valid[0, 0]) == 0:
        return None
    return patches[0, 0].cpu().numpy()


class DummyImageEncoder(object):
    def __init__(self):
        self.height, self.width = 16, 8
        self.image_shape = 16, 8, 3
        self.feature_dim = 128

    def __call__(self, data_in, batch_size=32):
        p = data_in if isinstance(data_in, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(data_in, np.uint8))
        out = ops.dummy_encode(p.cuda().contiguous())
        return out if isinstance(data_in, torch.Tensor) else out.cpu().numpy()


class ConstantImageEncoder(object):
    def __init__(self):
        self.height, self.width = 16, 8
        self.image_shape = 16, 8, 3
        self.feature_dim = 128

    def __call__(self, data_in, batch_size=32):
        if isinstance(data_in, torch.Tensor):
            out = torch.zeros((data_in.shape[0], 128), dtype=torch.float32, device=data_in.device)
            out[:, 0] = 1
            return out
        out = np.zeros((len(data_in), 128), dtype=np.float32)
        out[:, 0] = 1
        return out


def create_box_encoder(model_filename, input_name="images", output_name="features", batch_size=32, num_threads=1,
                       image_encoder=None):
    """generate_detections.py:180-215.  ``encoder(image, boxes, timing=False)`` for one frame, plus the additive
    ``encoder.batch(frames, boxes, counts)`` on device tensors for many frames at once."""
    if image_encoder is None:
        if 'dummy' in model_filename:
            image_encoder = DummyImageEncoder()
        elif 'constant' in model_filename:
            image_encoder = ConstantImageEncoder()
        else:
            raise NotImplementedError("CNN encoders are out of scope of deepdish_b200; pass image_encoder=")
    image_shape = tuple(int(v) for v in image_encoder.image_shape)

    def batch(frames, boxes, counts=None, boxes_are_int=True):
        """frames u8 [b,H,W,3], boxes f64 [b,dmax,4], counts i32 [b] (device tensors) ->
        (features [b,dmax,feature_dim], valid i32 [b,dmax])."""
        patches, valid = ops.extract_patches(frames, boxes, counts, image_shape[:2], boxes_are_int=boxes_are_int)
        b, dmax = valid.shape
        feats = image_encoder(patches.reshape((b * dmax,) + image_shape), batch_size)
        return feats.reshape(b, dmax, -1), valid

    def encoder(image, boxes, timing=False):
        if len(boxes) == 0:                                                # :190-194
            return (np.array([]), 0) if timing else np.array([])
        frame = torch.as_tensor(np.ascontiguousarray(image, dtype=np.uint8)).cuda()[None]
        bt, is_int = _boxes_tensor(boxes)
        feats, valid = batch(frame, bt, None, is_int)
        if int(valid.min()) == 0:     # the reference substitutes np.random noise for such a patch (:200-203)
            print("WARNING: Failed to extract image patch: %s." % str(np.asarray(boxes)[(valid[0] == 0).cpu().numpy()]))
        result = feats[0].cpu().numpy()
        return (result, 0.0) if timing else result

    encoder.image_encoder = image_encoder
    encoder.image_shape = image_shape
    encoder.width, encoder.height = image_encoder.width, image_encoder.height
    encoder.batch = batch
    return encoder
