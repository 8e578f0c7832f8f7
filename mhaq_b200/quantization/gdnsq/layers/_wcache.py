"""Per-step cache of the fake-quantized weight (BASELINE north_star (3)).

The reference re-quantizes every weight in every ``forward`` call
(gdnsq_conv2d.py:98).  A weight only changes when the optimizer steps, so:

* under ``torch.no_grad()`` (validation / calibration / prediction) the quantized
  weight is computed once and reused for every batch until a parameter changes;
* with autograd on, it is reused by further forward calls of the same layer made
  before the backward pass (weight-shared modules, several micro-batches whose
  losses are summed) — the graph is then traversed once with the accumulated
  gradient.  Running backward frees the graph, so the cache entry is dropped the
  moment its backward executes.

The key is built from tensor identity + in-place version counters, so optimizer
steps (`param.add_` bumps `_version`), `load_state_dict` (copy_) and calibration
(which REPLACES the Parameter objects, minmaxobserver.py:59-66,86) all invalidate
it.  Nothing is stored in the state dict.

One thing the key cannot see: a write through ``param.data`` (``p.data.copy_()``, old-style EMA /
weight clipping) does not bump ``param._version``.  With autograd on this does not matter (the
entry is dropped by the backward pass of every step); code that mutates ``.data`` between two
``no_grad`` forward calls must call ``layer._wq_cache.clear()`` (or use ``with torch.no_grad():
p.copy_()``, which does bump the version).
"""
from __future__ import annotations

import weakref

import torch


def _sig(t):
    if t is None:
        return None
    return (id(t), t._version, t.data_ptr(), tuple(t.shape))


class WeightQuantCache:
    __slots__ = ("key", "value", "hits", "misses", "_refs", "_probe")

    def __init__(self):
        self.key = None
        self.value = None
        self.hits = 0
        self.misses = 0
        self._refs = ()       # weak references to the tensors the entry was computed from
        self._probe = ()

    def lookup(self, tensors, grad_mode: bool, training: bool):
        key = (tuple(_sig(t) for t in tensors), grad_mode, training)
        self._probe = tensors
        if self.key == key and self.value is not None and len(self._refs) == len(tensors) and all(
                (r is None and t is None) or (r is not None and r() is t) for r, t in zip(self._refs, tensors)):
            # identity, not just id(): a re-bound Parameter (calibration replaces them) may reuse the
            # id, data_ptr and version 0 of a tensor that has been freed
            self.hits += 1
            return key, self.value
        self.misses += 1
        return key, None

    def store(self, key, value):
        self.key, self.value = key, value
        self._refs = tuple(None if t is None else weakref.ref(t) for t in self._probe)
        cache = self

        def _drop(grad, _key=key):
            # backward through this entry is running: its graph is about to be freed
            if cache.key == _key:
                cache.key, cache.value = None, None
            return grad

        for t in (value if isinstance(value, (tuple, list)) else (value,)):
            if torch.is_tensor(t) and t.requires_grad and t.grad_fn is not None:
                t.register_hook(_drop)

    def clear(self):
        self.key, self.value, self._refs = None, None, ()
