"""Quantizer plugins (reference: src/quantization/__init__.py:1-3).  ``DummyQuant`` — the
reference's no-op example plugin — has no arithmetic and is out of scope."""
from .gdnsq.gdnsq_quant import GDNSQQuant

__all__ = ["GDNSQQuant"]
