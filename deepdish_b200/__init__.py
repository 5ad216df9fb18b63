"""deepdish_b200 -- B200-native (sm_100a CUDA) tracking-by-detection hot path of AdaptiveCity/deepdish."""
__version__ = "0.1.0"
