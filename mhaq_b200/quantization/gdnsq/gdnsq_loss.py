"""Bit-width constraint losses — mirror of the reference's
src/quantization/gdnsq/gdnsq_loss.py:6-168 (``PotentialLoss`` with a prediction/target
pair, ``PotentialLossNoPred`` with a precomputed base loss).

    ploss = calib_mul * l1 * (wmul*wloss + amul*aloss) + l2 * rloss
    wloss = mean(max(0, (log_w_range - log_wght_s) - (w_target - eps))^p)   (same for act)

Row (f)-1 / (f)-4 of SURVEY.md §8.  On the GPU (fp32, p == 1) the whole expression and its autograd
are one kernel each way (`mhaq_fq_potential_loss_fwd/bwd_f32`) instead of ~30 tiny ATen launches;
otherwise the same operations in plain PyTorch."""
import torch
import torch.nn as nn


class _PotentialBase(nn.Module):
    def __init__(self, criterion, p=1, a=8, w=4, lossless=False) -> None:
        super().__init__()
        self.criterion = criterion
        self.s_weight_loss = torch.tensor(0)
        self.s_act_loss = torch.tensor(0)
        self.weight_reg_loss = torch.tensor(0)
        self.p = torch.tensor(p)
        self._p_int, self._eps = (int(p) if float(p) == int(p) else None), 1e-3   # host copies: no device sync per step
        self.at = a
        self.wt = w
        self.lossless = lossless
        self.l_eps = torch.tensor(1e-3)
        self.r_eps = torch.tensor(1e-3)
        self.aloss = torch.tensor(1.0)
        self.wloss = torch.tensor(1.0)
        self.loss_sum = 0.0
        self.cnt = 1
        self.t = 0.0

    def make_capturable(self, device):
        """Keep the running calibration multiplier state (`loss_sum`, `cnt`) in device tensors
        so that a CUDA-graph replay advances it (a Python int would be frozen at capture)."""
        self.loss_sum = torch.as_tensor(float(self.loss_sum), dtype=torch.float32, device=device).clone()
        self.cnt = torch.as_tensor(float(self.cnt), dtype=torch.float32, device=device).clone()
        # 0-dim CPU constants are legal operands of CUDA ops but their use in backward is a
        # host->device copy, which a graph capture forbids
        self.p, self.l_eps, self.r_eps = (t.to(device) for t in (self.p, self.l_eps, self.r_eps))
        return self

    def _fusable(self, base_loss, las, laq, lws, lwq):
        """One kernel each way when everything lives on the GPU in fp32 and p == 1 (the only
        exponent GDNSQQuant passes, gdnsq_quant.py:90-102)."""
        return (self._p_int == 1 and not torch.is_tensor(self.t) and all(torch.is_tensor(v) and v.is_cuda and v.dtype == torch.float32
                                         for v in (base_loss, las, laq, lws, lwq))
                and las.numel() == laq.numel() and lws.numel() == lwq.numel())

    def _combine_fused(self, base_loss, las, laq, lws, lwq):
        from ... import ops
        dev = lws.device
        if not (torch.is_tensor(self.loss_sum) and self.loss_sum.is_cuda and torch.is_tensor(self.cnt)
                and self.cnt.is_cuda):
            # the running calibration state lives in device scalars from the first GPU step on
            # (the reference's `loss_sum` becomes a device tensor at its first `+=` anyway)
            self.loss_sum = torch.as_tensor(float(self.loss_sum), dtype=torch.float32, device=dev).reshape(1).clone()
            self.cnt = torch.as_tensor(float(self.cnt), dtype=torch.float32, device=dev).reshape(1).clone()
        self.base_loss = base_loss
        ploss, rec = ops.potential_loss(base_loss, las, laq, lws, lwq, self.loss_sum, self.cnt, self.wt, self.at,
                                        self._eps, float(self.t), self.lossless, self.training)
        (self.wloss, self.aloss, self.rloss, self.s_weight_loss, self.q_weight_loss, self.s_act_loss,
         self.q_act_loss, self.weight_reg_loss) = rec[1:9].unbind(0)
        return ploss

    def _combine(self, base_loss, las, laq, lws, lwq):
        if self._fusable(base_loss, las, laq, lws, lwq):
            return self._combine_fused(base_loss, las, laq, lws, lwq)
        self.base_loss = base_loss
        z = torch.zeros((), dtype=torch.int64, device=lws.device)   # == torch.tensor(0), without a host copy
        wloss0 = torch.max(z, (lwq - lws) - (self.wt - self.l_eps)).pow(self.p)
        wloss = wloss0.mean()
        wact = (wloss0 > 0).sum()          # active weight constraints
        aloss0 = torch.max(z, (laq - las) - (self.at - self.l_eps)).pow(self.p)
        aloss = aloss0.mean()
        aact = (aloss0 > 0).sum()          # active activation constraints
        rloss = base_loss.pow_(self.p)
        calib_mul = self.loss_sum / self.cnt
        wmul = (wact + self.l_eps) / (wact + aact + self.l_eps)
        amul = (aact + self.l_eps) / (wact + aact + self.l_eps)
        l1, l2 = (1.0, self.t) if self.lossless else (self.t, 1.0)
        ploss = calib_mul * l1 * (wmul * wloss + amul * aloss) + l2 * rloss
        if self.training:
            self.loss_sum += rloss.detach()
            self.cnt += 1          # in place when `cnt` is a tensor (see make_capturable)
        self.wloss, self.aloss, self.rloss = wloss, aloss, rloss
        self.s_weight_loss = -lws.mean()
        self.q_weight_loss = lwq.mean()
        self.s_act_loss = -las.mean()
        self.q_act_loss = laq.mean()
        self.weight_reg_loss = (lwq - lws).max()
        return ploss


class PotentialLoss(_PotentialBase):
    def forward(self, output, target):
        """output = (prediction, log_act_s, log_act_q, log_wght_s, log_w)"""
        return self._combine(self.criterion(output[0], target), *output[1:5])


class PotentialLossNoPred(_PotentialBase):
    def forward(self, output):
        """output = (base_loss, log_act_s, log_act_q, log_wght_s, log_w)"""
        return self._combine(output[0], *output[1:5])
