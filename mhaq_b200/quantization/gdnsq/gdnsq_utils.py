"""Estimator / mode enumerations of the GDNSQ quantizer.

``QNMethod`` selects the gradient estimator of the rounding step; its integer values are the
reference's (src/quantization/gdnsq/gdnsq_utils.py:9-13) AND the `method` ids of the C ABI
(`MHAQ_FQ_STE / EWGS / AEWGS / LSQ` in include/mhaq_fq.h) — `tests/test_abi.py` pins the two
together.  ``QMode`` (gdnsq_utils.py:3-7) is vestigial in the reference — written, never read by
any arithmetic — and is kept only so that code importing it keeps importing.
"""
from enum import Enum

from ... import _lib  # noqa: F401  (the ABI the ids below must agree with)

_ESTIMATORS = ("STE", "EWGS", "AEWGS", "LSQ")            # index == ABI method id
_MODES = ("NOISE_VAL", "ROUND_VAL", "SOURCE_VAL", "FLOAT_TRAIN_VAL")

QNMethod = Enum("QNMethod", {n: i for i, n in enumerate(_ESTIMATORS)}, module=__name__, qualname="QNMethod")
QMode = Enum("QMode", {n: i + 1 for i, n in enumerate(_MODES)}, module=__name__, qualname="QMode")
