"""`quantization.params` block of a GDNSQ config — same keys and defaults as the reference's
src/quantization/gdnsq/config/config_schema.py:5-9, plus the checks the reference leaves to a
later `KeyError` (an unknown estimator name fails here, at config load)."""
from typing import Optional

from pydantic import BaseModel, field_validator

from ..gdnsq_utils import QNMethod


class GDNSQQuantizerParams(BaseModel):
    # knowledge distillation from a frozen copy of the FP model (gdnsq_quant.py:76-89)
    distillation: Optional[bool] = False
    distillation_loss: Optional[str] = "Cross-Entropy"
    distillation_teacher: Optional[str] = None
    # gradient estimator of the WEIGHT quantizers (activations always use STE, quirk 1)
    qnmethod: str = QNMethod.STE.name

    @field_validator("qnmethod")
    @classmethod
    def _known_estimator(cls, v: str) -> str:
        if v not in QNMethod.__members__:
            raise ValueError(f"qnmethod must be one of {sorted(QNMethod.__members__)}, got {v!r}")
        return v
