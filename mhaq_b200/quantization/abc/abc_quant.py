"""Plugin base class — mirror of the reference's src/quantization/abc/abc_quant.py:8-126.

A quantizer plugin exposes ``quantize(lmodel, in_place=False)`` plus the hooks
``module_mappings``, ``_quantize_module``, ``_get_quantization_sequence``,
``_get_layers`` and ``_init_config``; ``Quantizer(config)()`` instantiates the class
named by ``config.quantization.name``."""
from abc import ABC, abstractmethod
from typing import Dict, List

from torch import nn


class BaseQuant(ABC):
    def __init__(self, config):
        self.config = config
        self.act_bit: int
        self.weight_bit: int
        self.excluded: List
        self._init_config()

    @abstractmethod
    def module_mappings(self) -> Dict:
        """{source layer type: quantized layer type}"""

    @abstractmethod
    def quantize(self, model, in_place=False):
        """Returns the quantization-ready version of the (Lightning) module."""

    @abstractmethod
    def _quantize_module(self, module: nn.Module, *args, **kwargs) -> nn.Module:
        """Quantized counterpart of one layer."""

    @abstractmethod
    def _get_quantization_sequence(self, qmodule: nn.Module, *args, **kwargs) -> nn.Module:
        """Combine the quantized layer with its activation quantizer."""

    def _get_layers(self, model: nn.Module, exclude_layers: List[str] = []):
        """{name: type} of the quantizable layers, minus `exclude_layers`
        (AttributeError for a name that is not a quantizable layer) — abc_quant.py:89-114."""
        kinds = tuple(self.module_mappings().keys())
        layers = {n: type(m) for n, m in model.named_modules() if issubclass(type(m), kinds)}
        for name in exclude_layers:
            if name not in layers:
                raise AttributeError(f"Layer name {name} is not found in the model.")
            layers.pop(name)
        return layers

    def _init_config(self):
        if self.config:
            qc = self.config.quantization
            self.act_bit = qc.act_bit
            self.weight_bit = qc.weight_bit
            self.excluded_layers = qc.excluded_layers
