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

def _reference_root():
    """$DEEPDISH_REFERENCE, else an installed copy under baseline/_ref, else /root/reference."""
    env = os.environ.get("DEEPDISH_REFERENCE")
    if env:
        return env
    local = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
    if os.path.isdir(os.path.join(local, "deep_sort")):
        return local
    return "/root/reference"


REFERENCE_ROOT = _reference_root()


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


# ------------------------------------------------------------------------------------------------
# The reference's own counting code: Pipeline.process_results (deepdish.py:1035-1139) driven
# without camera / MQTT / web app.  Build container only.
# ------------------------------------------------------------------------------------------------
_pipeline_mod = None


def load_pipeline_module():
    """Load /root/reference/deepdish.py (not the deepdish/ package) with I/O dependencies stubbed."""
    global _pipeline_mod
    if _pipeline_mod is not None:
        return _pipeline_mod
    load()
    import importlib.util

    def stub(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class _Any:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, n):
            return _Any()

        def __call__(self, *a, **k):
            return _Any()

        def route(self, *a, **k):
            return lambda f: f

        before_serving = after_serving = route

    stub("cameratransform")
    stub("uvloop", install=lambda: None)
    stub("aiofiles")
    stub("gmqtt", Client=_Any)
    stub("quart", Quart=_Any, Response=_Any, current_app=_Any())
    stub("hypercorn")
    stub("hypercorn.asyncio", serve=lambda *a, **k: None)
    stub("hypercorn.config", Config=_Any)
    spec = importlib.util.spec_from_file_location("deepdish_main_ref", os.path.join(REFERENCE_ROOT, "deepdish.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _pipeline_mod = mod
    return mod


class RefCounter:
    """Runs the reference's Pipeline.process_results on a reference Tracker, one frame at a time."""

    def __init__(self, tracker, labels, line):
        import asyncio
        mod = load_pipeline_module()
        P = mod.Pipeline
        p = object.__new__(P)
        p.running = True
        p.args = types.SimpleNamespace(object_annotation="none", mqtt_verbosity=0)
        p.tracker = tracker
        from deepdish.framerecords import FrameRecords
        p.framerec = FrameRecords({})
        p.db = {}
        p.cameracountline = np.asarray(line, dtype=float).reshape(2, 2)
        p.trackdata_ratios = (1, 1)
        p.cam = None
        p.topdownview = None
        p.data_lock = asyncio.Lock()
        p.wanted_labels = list(labels)
        p.poscount = {l: 0 for l in labels}
        p.negcount = {l: 0 for l in labels}
        p.intcount = {l: 0 for l in labels}
        p.delcount = {l: 0 for l in labels}
        p.mqtt = None
        p.log = None
        p.cpu_temp_file = None
        p.output = None
        p.framenum_committed = 0
        self.p, self.mod, self.labels = p, mod, list(labels)
        self.frame = 0

    def step(self, detections):
        import asyncio
        import contextlib
        import io
        p = self.p

        class QIn:
            async def get(self_inner):
                return (self.frame, detections, [self.mod.FrameInfo(0.0, self.frame)], 0.0)

        class QOut:
            async def put(self_inner, item):
                p.running = False

        p.running = True
        with contextlib.redirect_stdout(io.StringIO()):
            asyncio.run(p.process_results(QIn(), QOut()))
        self.frame += 1

    def counts(self):
        p = self.p
        return np.array([[p.poscount[l], p.negcount[l], p.intcount[l], p.delcount[l]] for l in self.labels],
                        dtype=np.int64)


class RefBoxFilter:
    """Runs the reference's own pre-NMS box filter -- the loop inside the coroutine Pipeline.detect_objects
    (deepdish.py:887-982, filter at :941-960) -- on a detector result, without camera, CNN or encoder: the queues,
    the executor and the detector are stubs, background subtraction is off (--disable-background-subtraction)."""

    def __init__(self, frame_w=640, frame_h=480):
        mod = load_pipeline_module()
        p = object.__new__(mod.Pipeline)
        p.running = True
        p.input_size = (frame_w, frame_h)
        p.frame_count = 0
        p.everyframe = None
        p.background_subtraction = False
        p.args = types.SimpleNamespace(enable_background_masking=False, object_detector_skip_frames=0,
                                       disable_powersaving=True, background_subtraction_ratio=0.0)
        p.powersave_delay = 0
        p.cameracountline = np.array([[frame_w / 2, 0], [frame_w / 2, frame_h]], dtype=float)
        p.pipeline_sem = types.SimpleNamespace(release=lambda: None)
        p.kickstart = types.SimpleNamespace(set=lambda: None)
        enc = lambda frame, boxes: None               # noqa: E731  (the encoder is only warmed up here)
        enc.width, enc.height = 64, 128
        p.encoder = enc
        self.p, self.mod = p, mod
        self.frame = 0

    def __call__(self, boxes0, labels0, scores0):
        """-> (boxes [(x, y, w, h) ints], labels, scores) exactly as detect_objects puts them on its output queue."""
        import asyncio
        p = self.p
        result = {}
        outer = self

        async def run_in_executor(_ex, fn, *a):
            if getattr(fn, "__name__", "") == "run_object_detector":
                return (boxes0, labels0, scores0, 0.0)
            return None                                   # the encoder warm-up call

        class Loop:
            pass

        loop = Loop()
        loop.run_in_executor = run_in_executor
        p.loop = loop

        class QIn:
            async def get(self_inner):
                outer.frame += 1
                return (outer.frame, np.zeros((4, 4, 4), np.uint8), 0.0, 0.0, 0.0)

        class QOut:
            async def put(self_inner, item):
                result["boxes"], result["labels"], result["scores"] = item[2], item[3], item[4]
                p.running = False

        p.running = True
        asyncio.run(p.detect_objects(QIn(), QOut()))
        return result["boxes"], result["labels"], result["scores"]
