"""Enums of the reference's src/quantization/gdnsq/gdnsq_utils.py:3-13.
QNMethod values are also the `method` ids of the C ABI (include/mhaq_fq.h)."""
from enum import Enum


class QMode(Enum):  # vestigial in the reference (never read by any arithmetic)
    NOISE_VAL = 1
    ROUND_VAL = 2
    SOURCE_VAL = 3
    FLOAT_TRAIN_VAL = 4


class QNMethod(Enum):
    STE = 0
    EWGS = 1
    AEWGS = 2
    LSQ = 3
