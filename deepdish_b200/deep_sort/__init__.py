"""Drop-in mirror of the reference's ``deep_sort`` package (same module / class / function names,
SURVEY.md section 8b) whose arithmetic runs in hand-written sm_100a CUDA kernels through the C ABI.

    import deepdish_b200; deepdish_b200.install_as_deep_sort()   # makes `import deep_sort` resolve here
"""
from . import detection, kalman_filter, nn_matching, iou_matching, linear_assignment, preprocessing, track, tracker  # noqa: F401
