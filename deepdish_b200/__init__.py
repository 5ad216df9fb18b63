"""deepdish_b200 -- B200-native (sm_100a CUDA) tracking-by-detection hot path of AdaptiveCity/deepdish."""
__version__ = "0.1.0"


def install_as_deep_sort(box_encoder=False, framerecords=False):
    """Make ``import deep_sort`` / ``from tools.intersection import ...`` (the imports of the reference's
    deepdish.py:49-57) resolve to this package's CUDA-backed mirror.  Raises if the CUDA library is missing.

    box_encoder=True also routes ``from tools import generate_detections`` (deepdish.py:55) here: patch extraction and
    the reference's arithmetic encoders (``--encoder-model dummy`` / ``constant``) then run on the GPU; CNN encoders are
    out of scope, so leave it False when a real MARS model is used.  framerecords=True routes
    ``deepdish.framerecords`` (deepdish.py:57) to the mirror -- not needed with the reference tree on the path: its own
    FrameRecords works unchanged on top of this Tracker, which writes the hooks' host edits back to the device."""
    import sys
    from . import _lib, deep_sort, tools
    from .tools import intersection
    _lib.lib()
    if framerecords:
        import types
        from . import framerecords as fr
        pkg = sys.modules.setdefault("deepdish", types.ModuleType("deepdish"))
        pkg.framerecords = fr
        sys.modules["deepdish.framerecords"] = fr
    sys.modules["deep_sort"] = deep_sort
    for name in ("detection", "kalman_filter", "nn_matching", "iou_matching", "linear_assignment",
                 "preprocessing", "track", "tracker"):
        sys.modules["deep_sort." + name] = getattr(deep_sort, name)
    sys.modules.setdefault("tools", tools)
    sys.modules["tools.intersection"] = intersection
    if box_encoder:
        from .tools import generate_detections
        sys.modules["tools.generate_detections"] = generate_detections
        sys.modules["tools"].generate_detections = generate_detections
