"""deepdish_b200 -- B200-native (sm_100a CUDA) tracking-by-detection hot path of AdaptiveCity/deepdish."""
__version__ = "0.1.0"


def install_as_deep_sort():
    """Make ``import deep_sort`` / ``from tools.intersection import ...`` (the imports of the reference's
    deepdish.py:49-57) resolve to this package's CUDA-backed mirror.  Raises if the CUDA library is missing."""
    import sys
    from . import _lib, deep_sort, tools
    from .tools import intersection
    _lib.lib()
    sys.modules["deep_sort"] = deep_sort
    for name in ("detection", "kalman_filter", "nn_matching", "iou_matching", "linear_assignment",
                 "preprocessing", "track", "tracker"):
        sys.modules["deep_sort." + name] = getattr(deep_sort, name)
    sys.modules.setdefault("tools", tools)
    sys.modules["tools.intersection"] = intersection
