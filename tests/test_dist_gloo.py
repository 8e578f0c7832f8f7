"""CPU, world_size 2 over gloo: the N>1 host logic.

* the packed AEWGS statistics all-reduce equals the reference's three all-reduces;
* the oracle's 2-rank AEWGS backward equals the closed form with rank-averaged statistics
  (i.e. what the kernels compute from the packed buffer);
* bench.py's reference arm under torchrun: rank 0 alone prints one JSON line."""
import json
import os
import socket
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        from mhaq_b200 import ops
        from oracle import fq_oracle as O
        torch.manual_seed(100 + rank)
        C = 6
        # (1) packed == three separate collectives
        num, e2, me = torch.randn(C), torch.rand(C), torch.randn(C) * 0.1
        packed = ops.allreduce_packed_stats(torch.cat([num, e2, me]).clone())
        for t in (num, e2, me):
            dist.all_reduce(t, op=dist.ReduceOp.AVG)
        ok1 = torch.equal(packed, torch.cat([num, e2, me]))
        # (2) oracle AEWGS under 2 ranks: replicated weights, per-rank upstream gradient
        g = torch.Generator().manual_seed(7)
        w = torch.randn(C, 4, 3, 3, generator=g) * 0.3
        log_s = torch.full((C, 1, 1, 1), -3.0)
        go = torch.randn(C, 4, 3, 3, generator=torch.Generator().manual_seed(50 + rank))
        r = torch.zeros_like(w)
        wr = w.clone().requires_grad_(True)
        ls = log_s.clone().requires_grad_(True)
        O.weight_fake_quant(wr, ls, True, "AEWGS", noise=r).backward(go)
        # closed form with rank-averaged statistics
        s = torch.exp2(log_s)
        zp = w.amin((1, 2, 3), keepdim=True)
        v = (w - zp) / s
        e = torch.round(v) - v
        gg = go * s
        stats = torch.stack([(gg.sign() * e).mean((1, 2, 3)), e.square().mean((1, 2, 3)),
                             e.mean((1, 2, 3))]).reshape(-1)
        stats = ops.allreduce_packed_stats(stats).reshape(3, C, 1, 1, 1)
        delta = stats[0] / (stats[1] - stats[2].square()).clamp_min(1e-3)
        gs = (delta * (gg.sign() * e)).clamp_max(0.99)
        gv = gg + (-gg * gs)
        gu = gv / s
        g_zp = (go - gu).sum((1, 2, 3), keepdim=True)
        expect = gu + torch.where(w == zp, g_zp / (w == zp).sum((1, 2, 3), keepdim=True), torch.zeros(()))
        ok2 = torch.allclose(wr.grad, expect, rtol=1e-5, atol=1e-6)
        q.put((rank, bool(ok1), bool(ok2), float((wr.grad - expect).abs().max())))
    finally:
        dist.destroy_process_group()


def test_packed_stats_allreduce_and_two_rank_aewgs():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok1, ok2, err in res:
        assert ok1, f"rank {rank}: packed all-reduce differs from three all-reduces"
        assert ok2, f"rank {rank}: 2-rank AEWGS oracle vs closed form, max err {err}"


def test_bench_reference_arm_under_torchrun_two_ranks():
    port = _free_port()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"),
           "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--cpu-log2n", "18"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["unit"] == "GB/s"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["steps"] == 1 and d["warmup"] == 1          # the arm runs exactly the steps it was asked for
    assert d["product_modules_loaded"] == []             # nothing of the product in the reference process


class _ToyQuantLayer(torch.nn.Module):
    """Parameter layout of a per-channel NoisyConv2d as DDP sees it: used weight + scale, and a
    trainable `log_b_s` that never takes part in the forward."""

    def __init__(self):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.randn(4, 3))
        self.log_wght_s = torch.nn.Parameter(torch.full((4, 1), -2.0))
        self.log_b_s = torch.nn.Parameter(torch.full((1,), -12.0))

    def forward(self, x):
        return x @ (self.weight * torch.exp2(self.log_wght_s)).t()


def _ddp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        from mhaq_b200 import harness
        torch.manual_seed(0)                       # identical replicas
        model = torch.nn.Sequential(_ToyQuantLayer(), torch.nn.ReLU(), torch.nn.Linear(4, 2))
        ddp = harness.wrap_ddp(model, None, lean=True)
        opt = torch.optim.SGD([p for p in model.parameters()], lr=0.1)
        g = torch.Generator().manual_seed(10 + rank)     # different data per rank
        for _ in range(3):                         # would raise on step 2 if log_b_s were not ignored
            x = torch.randn(8, 3, generator=g)
            ddp(x).square().mean().backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
        # the replicas-in-sync check of bench.py's QAT leg
        chk = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        unused_grad_none = model[0].log_b_s.grad is None
        q.put((rank, bool((lo == hi).item()), unused_grad_none, float(model[0].log_b_s.detach())))
    finally:
        dist.destroy_process_group()


def test_lean_ddp_wrapping_ignores_the_never_used_log_b_s():
    """harness.wrap_ddp(lean=True): no find_unused_parameters graph walk — the unused `log_b_s`
    parameters are excluded from DDP's reducer instead; replicas stay in sync."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, in_sync, unused_none, lbs in res:
        assert in_sync, f"rank {rank}: replicas diverged"
        assert unused_none and lbs == -12.0
