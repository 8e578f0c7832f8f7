"""Enums of the reference's src/aux/types.py:3-25 (names and values are part of the
config / plugin contract; only QScheme is read on the hot path)."""
from enum import Enum


class DType(Enum):
    VISION_CLS = 1
    VISION_SR = 2
    VISION_DNS = 3
    VISION_OD = 4


class MType(Enum):
    VISION_CLS = 1
    VISION_SR = 2
    VISION_DNS = 3
    VISION_OD = 4
    LM = 10


class QScheme(Enum):
    PER_TENSOR = 0
    PER_CHANNEL = 1


class QMethod(Enum):
    GDNSQ = 0
