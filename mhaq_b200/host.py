"""Fake-quant forward + backward for tensors that live in HOST memory.

``fake_quant_fwd_bwd_host`` streams a host-resident tensor through the GPU in row chunks:
while chunk *i* is being quantized (one forward and one backward kernel), chunk *i+1* is on
its way in over PCIe and the results of chunk *i-1* are on their way out, on three CUDA
streams with `stages`-deep device staging (3 by default: one chunk arriving, one computing or
waiting, one leaving, so a late copy in one direction does not stall the other).  PCIe is full duplex, so the end-to-end time
approaches max(bytes in, bytes out) / link bandwidth instead of their sum — this is the
`e2e` leg of ``bench.py`` (inputs and outputs in pinned host memory, copies inside the timed
region).

Chunks are whole rows, so per-channel parameter gradients need no cross-chunk reduction
(rows are channels); for a per-tensor quantizer the per-chunk gradients are summed on the
device in chunk order (deterministic).
"""
from __future__ import annotations

import math

import torch

from . import ops


def _rows_view(t: torch.Tensor, rows: int) -> torch.Tensor:
    return t.reshape(rows, -1)


def fake_quant_fwd_bwd_host(x_host: torch.Tensor, go_host: torch.Tensor, scale, zero_point,
                            min_val=None, max_val=None, method="STE", y_host=None, gx_host=None,
                            chunks: int = 8, device=None, philox=(0, 0), stages: int = 3):
    """x_host, go_host: pinned fp32 host tensors of identical shape.  Parameters are CUDA
    tensors (per-tensor: numel 1; per-channel: one per row of dim 0).  Returns
    (y_host, gx_host, grads) with grads = dict(scale=..., zero_point=..., min_val=..., max_val=...)
    as CUDA tensors shaped like the parameters.  The call returns after all device work has
    been enqueued and the output copies issued; synchronise before reading the host outputs."""
    device = torch.device(device) if device is not None else scale.device
    if ops._method_id(method) == ops.METHOD_IDS["AEWGS"]:
        # AEWGS needs whole-tensor (per-channel) statistics — all-reduced under DDP — BEFORE any
        # chunk's backward can run: it cannot be streamed chunk by chunk.  Fail before any copy.
        raise NotImplementedError("fake_quant_fwd_bwd_host streams the tensor in chunks; the AEWGS estimator "
                                  "needs its per-channel statistics over the whole tensor first — use "
                                  "mhaq_b200.fake_quant on a device-resident tensor")
    assert x_host.shape == go_host.shape and x_host.dtype == torch.float32
    if not (x_host.is_pinned() and go_host.is_pinned()):
        raise RuntimeError("fake_quant_fwd_bwd_host needs pinned host tensors")
    rows = x_host.shape[0] if x_host.dim() > 1 else 1
    per_channel = scale.numel() > 1
    if per_channel and scale.numel() != rows:
        raise RuntimeError("per-channel parameters must have one entry per row of dim 0")
    x2, g2 = _rows_view(x_host, rows), _rows_view(go_host, rows)
    inner = x2.shape[1]
    if rows == 1:                                   # per-tensor vector: chunk along the inner dim
        n_chunks = max(1, min(chunks, inner // (1 << 16)))
        cuts = [(0, 1, i * (inner // n_chunks), inner if i == n_chunks - 1 else (i + 1) * (inner // n_chunks))
                for i in range(n_chunks)]
    else:
        n_chunks = max(1, min(chunks, rows))
        step = (rows + n_chunks - 1) // n_chunks
        cuts = [(r0, min(rows, r0 + step), 0, inner) for r0 in range(0, rows, step)]
    if y_host is None:
        y_host = torch.empty(x_host.shape, dtype=torch.float32).pin_memory()
    if gx_host is None:
        gx_host = torch.empty(x_host.shape, dtype=torch.float32).pin_memory()
    y2, gx2 = _rows_view(y_host, rows), _rows_view(gx_host, rows)

    mid = ops._method_id(method)
    cur = torch.cuda.current_stream(device)
    s_in, s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)
    s_in.wait_stream(cur)
    max_elems = max((r1 - r0) * (c1 - c0) for r0, r1, c0, c1 in cuts)
    stages = max(2, min(stages, len(cuts)))
    stage = [[torch.empty(max_elems, dtype=torch.float32, device=device) for _ in range(4)]
             for _ in range(stages)]               # [x, go, y, gx] x `stages` buffers in flight
    ev_in = [torch.cuda.Event() for _ in cuts]
    ev_done = [torch.cuda.Event() for _ in cuts]
    ev_out = [None] * stages
    pshape = None
    acc = {k: None for k in ("scale", "zero_point", "min_val", "max_val")}
    prm = dict(scale=scale, zero_point=zero_point, min_val=min_val, max_val=max_val)

    def sub(p, r0, r1):
        if not torch.is_tensor(p) or p.numel() == 1:
            return p
        return p.reshape(rows, -1)[r0:r1].reshape((r1 - r0,) + (1,) * (1 if rows > 1 else 0))

    def issue_in(i):
        r0, r1, c0, c1 = cuts[i]
        n = (r1 - r0) * (c1 - c0)
        b = stage[i % stages]
        with torch.cuda.stream(s_in):
            if ev_out[i % stages] is not None:
                s_in.wait_event(ev_out[i % stages])      # staging buffers free again
            b[0][:n].view(r1 - r0, c1 - c0).copy_(x2[r0:r1, c0:c1], non_blocking=True)
            b[1][:n].view(r1 - r0, c1 - c0).copy_(g2[r0:r1, c0:c1], non_blocking=True)
            ev_in[i].record(s_in)

    issue_in(0)
    for i, (r0, r1, c0, c1) in enumerate(cuts):
        if i + 1 < len(cuts):
            issue_in(i + 1)
        n = (r1 - r0) * (c1 - c0)
        b = stage[i % stages]
        shape = (r1 - r0, c1 - c0)
        cur.wait_event(ev_in[i])
        xd, gd = b[0][:n].view(shape), b[1][:n].view(shape)
        sc, zp, lo, hi = (sub(prm[k], r0, r1) for k in ("scale", "zero_point", "min_val", "max_val"))
        L = ops._Launch(xd, sc, zp, -math.inf if lo is None else lo, math.inf if hi is None else hi)
        geo = L.geo
        yd, gxd = b[2][:n].view(shape), b[3][:n].view(shape)
        ops.check(ops.lib.mhaq_fq_fwd_f32(xd.data_ptr(), yd.data_ptr(), None, *L.params(), geo.n_rows,
                                          geo.n_inner, geo.n_ch, None, ops._stream()), "mhaq_fq_fwd_f32")
        ws = ops._workspace(xd, geo)
        tk = ops._tickets(xd, geo)
        out = torch.empty(4, geo.n_ch, dtype=torch.float32, device=device)
        ops.check(ops.lib.mhaq_fq_bwd_f32(gd.data_ptr(), xd.data_ptr(), gxd.data_ptr(), *L.params(),
                                          geo.n_rows, geo.n_inner, geo.n_ch, mid, 0, None,
                                          philox[0], philox[1] + i, None, None, ws.data_ptr(),
                                          ops._stream()), "mhaq_fq_bwd_f32")
        ops.check(ops.lib.mhaq_fq_bwd_finalize_f32(ws.data_ptr(), tk.data_ptr(), *L.params(), geo.n_rows,
                                                   geo.n_inner, geo.n_ch, out[0].data_ptr(),
                                                   out[1].data_ptr(), out[2].data_ptr(),
                                                   out[3].data_ptr(), ops._stream()),
                  "mhaq_fq_bwd_finalize_f32")
        ev_done[i].record(cur)
        for j, k in enumerate(("scale", "zero_point", "min_val", "max_val")):
            p = prm[k]
            if not torch.is_tensor(p):
                continue
            if p.numel() == 1:                       # per-tensor parameter: sum over chunks/channels
                v = out[j].sum().reshape(p.shape)
                acc[k] = v if acc[k] is None else acc[k] + v
            else:
                if acc[k] is None:
                    acc[k] = torch.empty(rows, dtype=torch.float32, device=device)
                acc[k][r0:r1] = out[j]
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_done[i])
            y2[r0:r1, c0:c1].copy_(yd, non_blocking=True)
            gx2[r0:r1, c0:c1].copy_(gxd, non_blocking=True)
            e = torch.cuda.Event()
            e.record(s_out)
            ev_out[i % stages] = e
        for t in (ws, out):
            t.record_stream(cur)
    cur.wait_stream(s_out)
    grads = {k: (None if v is None else v.reshape(prm[k].shape)) for k, v in acc.items()}
    return y_host, gx_host, grads
