"""B200-native (sm_100a) implementation of the SegFormer teacher-student hot path of
JiaqiLi404/SemiSupervisedObjectDetection.  Importing this package never touches CUDA; the first kernel call loads
libsfk.so and fails loudly if it or an sm_100 device is missing (there is no fallback path)."""
from .configuration import SegformerConfig  # noqa: F401

__all__ = ["SegformerConfig"]
