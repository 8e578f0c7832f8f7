"""Task / model / quantization-scheme enumerations.

Names and integer values are the config and plugin contract of the reference
(src/aux/types.py:3-25): YAML configs carry the integers (`qscheme: 1`), `config_schema`
validates against the names.  Only ``QScheme`` is read on the hot path.  Declared from tables so
that the contract is visible in one place; members behave exactly like class-syntax ``Enum``s
(`QScheme.PER_CHANNEL.value == 1`, `QScheme["PER_TENSOR"]`, `QScheme(0)`, picklable).
"""
from enum import Enum

_VISION_TASKS = {"VISION_CLS": 1, "VISION_SR": 2, "VISION_DNS": 3, "VISION_OD": 4}


def _enum(name, members):
    return Enum(name, members, module=__name__, qualname=name)


#: dataset families
DType = _enum("DType", _VISION_TASKS)
#: model families: the vision tasks plus language models
MType = _enum("MType", {**_VISION_TASKS, "LM": 10})
#: granularity of the weight scale / zero point: one per tensor, or one per output channel
QScheme = _enum("QScheme", {"PER_TENSOR": 0, "PER_CHANNEL": 1})
#: quantization algorithm family
QMethod = _enum("QMethod", {"GDNSQ": 0})


def scheme_id(qscheme) -> int:
    """0 = per tensor, 1 = per channel, from this package's ``QScheme``, the integer a YAML config
    carries, the member name, or ANOTHER package's enum with the same contract — the reference's
    own ``src.aux.types.QScheme`` when its ``GDNSQQuant`` constructs this repo's layer classes
    (INTEGRATION.md §B).  Stored attributes are kept as given; comparisons go through here."""
    if isinstance(qscheme, QScheme):
        return qscheme.value
    if isinstance(qscheme, Enum):
        name, value = qscheme.name, qscheme.value
        if name in QScheme.__members__ and QScheme[name].value == value:
            return value
        raise ValueError(f"unknown quantization scheme {qscheme!r}")
    if isinstance(qscheme, str):
        return QScheme[qscheme].value
    return QScheme(int(qscheme)).value


def is_per_channel(qscheme) -> bool:
    return scheme_id(qscheme) == QScheme.PER_CHANNEL.value


def is_per_tensor(qscheme) -> bool:
    return scheme_id(qscheme) == QScheme.PER_TENSOR.value
