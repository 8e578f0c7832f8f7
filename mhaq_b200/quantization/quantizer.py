"""Plugin factory — mirror of the reference's src/quantization/quantizer.py:6-12:
``Quantizer(config)()`` returns an instance of the class named by
``config.quantization.name`` looked up on this package."""
from typing import Any


class Quantizer:
    def __init__(self, config) -> None:
        self.config = config

    def __call__(self) -> Any:
        from .. import quantization as compose_quantization
        return getattr(compose_quantization, self.config.quantization.name)(self.config)
