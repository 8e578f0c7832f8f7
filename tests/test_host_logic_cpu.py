"""CPU: host-side layout logic of mhaq_b200/ops.py (no kernel is launched) — geometry inference,
which memory layouts are walked in place, and the [rows, inner] storage-order views."""
import pytest
import torch

from mhaq_b200 import ops


def test_infer_geometry_per_tensor_and_per_channel():
    x = torch.empty(8, 4, 3, 3)
    g = ops.infer_geometry(x, [torch.empty(1)])
    assert (g.n_rows, g.n_inner, g.n_ch, g.axis) == (1, 288, 1, None)
    g = ops.infer_geometry(x, [torch.empty(8, 1, 1, 1), torch.empty(1)])
    assert (g.n_rows, g.n_inner, g.n_ch, g.axis) == (8, 36, 8, 0)
    assert g.param_shape(x.shape) == (8, 1, 1, 1)
    # a channel axis that is not dim 0: rows = outer x channel, channel = row % n_ch
    g = ops.infer_geometry(torch.empty(2, 5, 7), [torch.empty(1, 5, 1)])
    assert (g.n_rows, g.n_inner, g.n_ch, g.axis) == (10, 7, 5, 1)
    # bias-style: value (O,), scale (O,)
    g = ops.infer_geometry(torch.empty(16), [torch.empty(16)])
    assert (g.n_rows, g.n_inner, g.n_ch, g.axis) == (16, 1, 16, 0)
    with pytest.raises(RuntimeError, match="one broadcast"):
        ops.infer_geometry(x, [torch.empty(8, 4, 1, 1)])
    with pytest.raises(RuntimeError, match="disagree"):
        ops.infer_geometry(x, [torch.empty(8, 1, 1, 1), torch.empty(1, 4, 1, 1)])
    with pytest.raises(RuntimeError, match="does not broadcast"):
        ops.infer_geometry(torch.empty(4, 4), [torch.empty(2, 2, 4, 4)])


def test_which_layouts_are_walked_in_place():
    x = torch.randn(2, 6, 5, 4)
    cl = x.contiguous(memory_format=torch.channels_last)
    assert ops._is_dense(x, None) and ops._is_dense(x, 0) and ops._is_dense(x, 1)
    assert ops._is_dense(cl, None) and ops._is_dense(cl, 0)
    assert not ops._is_dense(cl, 1)                    # per-channel along C is not a dense row in NHWC
    assert ops._dense(cl, None) is cl                  # no copy
    assert ops._dense(cl, 1).is_contiguous()           # copy to row-major
    x5 = torch.randn(2, 3, 4, 5, 6).contiguous(memory_format=torch.channels_last_3d)
    assert ops._is_dense(x5, None) and ops._is_dense(x5, 0)
    assert not ops._is_dense(x.transpose(0, 1), None)  # an arbitrary permutation is copied
    assert not ops._is_dense(x[:, ::2], None)          # so is a strided slice
    assert ops._is_dense(torch.randn(7, 3).t().contiguous(), 0)


def test_rows_view_round_trip_in_storage_order():
    w = torch.randn(6, 4, 3, 3)
    for t in (w, w.contiguous(memory_format=torch.channels_last)):
        r = ops._rows2d(t)
        assert r.shape == (6, 36) and r.is_contiguous() and r.data_ptr() == t.data_ptr()
        # each row of the view is exactly the elements of one dim-0 slice (in storage order)
        for i in range(6):
            assert torch.equal(r[i].sort().values, t[i].reshape(-1).sort().values)
        back = ops._unrows(r, t)
        assert back.shape == t.shape and back.stride() == t.stride() and torch.equal(back, t)
    # layout transfer only when the strides differ
    a = torch.randn(2, 3, 4, 5)
    b = a.contiguous(memory_format=torch.channels_last)
    assert ops._like_layout(a, a) is a
    moved = ops._like_layout(a, b)
    assert moved.stride() == b.stride() and torch.equal(moved, a)


def test_fusable_weight_rows_policy_and_stats_quirk_flag():
    w = torch.empty(8, 4, 3, 3)
    ls = torch.empty(8, 1, 1, 1)
    assert ops.weight_rows_fusable(w, ls, "STE") and ops.weight_rows_fusable(w, ls, "LSQ")
    assert ops.weight_rows_fusable(w, ls, "EWGS") and not ops.weight_rows_fusable(w, ls, "AEWGS")
    assert not ops.weight_rows_fusable(w, torch.empty(1), "STE")             # per-tensor scale
    assert not ops.weight_rows_fusable(torch.empty(2, ops.WROW_MAX_INNER + 4), torch.empty(2, 1), "STE")
    assert ops.weight_rows_fusable(torch.empty(2, ops.WROW_MAX_INNER), torch.empty(2, 1), "STE")
    assert ops._method_id("AEWGS") == 2 and ops._method_id(3) == 3
    with pytest.raises(AttributeError, match="Unknown method"):
        ops._method_id(7)


def test_cpu_tensors_are_refused_everywhere():
    x = torch.randn(4, 4)
    s = torch.ones(1)
    for call in (lambda: ops.fake_quant(x, s, s),
                 lambda: ops.quantize_codes(x, s, s),
                 lambda: ops.act_fake_quant(x, s, s, s),
                 lambda: ops.weight_fake_quant_rows(x, torch.ones(4, 1)),
                 lambda: ops.weight_fake_quant_log(x, torch.ones(4, 1)),
                 lambda: ops.row_stats(x)):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()
    with pytest.raises(RuntimeError, match="fp32 only"):
        ops._require_cuda(_FakeCuda())


class _FakeCuda:
    """Stands in for a CUDA half tensor in the dtype check (no GPU in the CPU suite)."""
    is_cuda = True
    dtype = torch.float16
