"""Helpers of the reference's src/aux/qutils.py:1-22 used by the model surgery."""


def attrsetter(*items):
    """Returns f(obj, val) that assigns `val` to each dotted attribute path in `items`
    (for a path ending in 'bias', assigns val.bias) — qutils.py:1-19."""

    def resolve(obj, path):
        *head, tail = path.split(".")
        for name in head:
            obj = getattr(obj, name)
        return obj, tail

    def setter(obj, val):
        for path in items:
            owner, attr = resolve(obj, path)
            setattr(owner, attr, val.bias if attr == "bias" else val)

    return setter


def is_biased(module) -> bool:
    return getattr(module, "bias", None) is not None
