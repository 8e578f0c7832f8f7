"""`quantization.params` schema — mirror of the reference's
src/quantization/gdnsq/config/config_schema.py:5-9."""
from typing import Optional

from pydantic import BaseModel


class GDNSQQuantizerParams(BaseModel):
    distillation: Optional[bool] = False
    distillation_loss: Optional[str] = "Cross-Entropy"
    distillation_teacher: Optional[str] = None
    qnmethod: str = "STE"
