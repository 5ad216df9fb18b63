"""Import the UNMODIFIED reference (``/root/reference``) with the shims it needs on numpy 2 / PIL 10.

TEST INFRASTRUCTURE (build container only -- ``/root/reference`` does not exist on the GPU box;
nothing in ``-m gpu`` tests, ``smoke()`` or ``bench.py`` imports this module).

Shims (applied before import, no reference file is edited):
  1. ``np.float = float; np.int = int``  -- ``deep_sort/detection.py:30``, ``deep_sort/preprocessing.py:40``
     use aliases removed in numpy >= 1.24.
  2. ``PIL.Image.ANTIALIAS = Image.LANCZOS`` -- ``tools/yolov5.py:99``, ``tools/ssd_mobilenet.py:55``.
  3. stub ``tflite_runtime.interpreter`` with a fake ``Interpreter`` that serves a synthetic head.
"""
import importlib
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("DEEPDISH_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "deep_sort"))


class FakeInterpreter:
    """Stands in for ``tflite_runtime.interpreter.Interpreter``; serves tensors set by the test."""
    input_shape = (1, 640, 640, 3)
    outputs = None          # list of np arrays served by get_tensor(index)

    def __init__(self, *a, **kw):
        pass

    def allocate_tensors(self):
        pass

    def get_input_details(self):
        return [{"shape": np.array(type(self).input_shape), "index": 1000,
                 "quantization": (1.0, 0)}]

    def get_output_details(self):
        return [{"index": i, "quantization": (1.0, 0)} for i in range(len(type(self).outputs or [0]))]

    def set_tensor(self, index, data):
        pass

    def invoke(self):
        pass

    def get_tensor(self, index):
        return type(self).outputs[index]


_loaded = {}


def load():
    """Return a namespace with the reference modules (deep_sort.*, tools.intersection, ...)."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "int"):
        np.int = int
    try:
        from PIL import Image
        if not hasattr(Image, "ANTIALIAS"):
            Image.ANTIALIAS = Image.LANCZOS
    except Exception:
        pass
    tfl = types.ModuleType("tflite_runtime")
    tfl_i = types.ModuleType("tflite_runtime.interpreter")
    tfl_i.Interpreter = FakeInterpreter
    tfl_i.load_delegate = lambda *a, **k: None
    tfl.interpreter = tfl_i
    sys.modules.setdefault("tflite_runtime", tfl)
    sys.modules.setdefault("tflite_runtime.interpreter", tfl_i)
    if "cv2" not in sys.modules:
        try:
            import cv2  # noqa: F401  (deep_sort/preprocessing.py:3 imports it, never uses it)
        except Exception:
            sys.modules["cv2"] = types.ModuleType("cv2")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    names = ["deep_sort.detection", "deep_sort.kalman_filter", "deep_sort.nn_matching",
             "deep_sort.iou_matching", "deep_sort.linear_assignment", "deep_sort.preprocessing",
             "deep_sort.track", "deep_sort.tracker", "tools.intersection"]
    for n in names:
        _loaded[n.replace(".", "_")] = importlib.import_module(n)
    for n in ["tools.yolov5", "tools.ssd_mobilenet"]:
        try:
            _loaded[n.replace(".", "_")] = importlib.import_module(n)
        except Exception as e:  # PIL / yaml missing -> detector adapters unavailable
            _loaded[n.replace(".", "_")] = None
            _loaded[n.replace(".", "_") + "_error"] = repr(e)
    return types.SimpleNamespace(**_loaded)
