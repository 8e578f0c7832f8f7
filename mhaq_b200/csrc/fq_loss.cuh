// fq_loss.cuh — PotentialLoss's bit-width constraint (reference src/quantization/gdnsq/
// gdnsq_loss.py:32-86 / 114-168, exponent p = 1 as GDNSQQuant always passes, gdnsq_quant.py:90-102)
// as ONE launch forward and ONE backward instead of ~30 tiny ATen launches per step.
//
//   wloss0_i = max(0, (lwq_i - lws_i) - (wt - eps))        wloss = mean_i wloss0_i   wact = #{wloss0_i > 0}
//   aloss0_j = max(0, (laq_j - las_j) - (at - eps))        aloss = mean_j aloss0_j   aact = #{aloss0_j > 0}
//   wmul = (wact + eps) / (wact + aact + eps)              amul = (aact + eps) / (wact + aact + eps)
//   ploss = calib * l1 * (wmul * wloss + amul * aloss) + l2 * rloss,   calib = loss_sum / cnt
//   training: loss_sum += rloss, cnt += 1
//
// Inputs are O(#channels) vectors (ResNet-18: 3904 weight channels, 16 activations): one CTA,
// fp64 accumulation.  Included by mhaq_fq.cu (inside its anonymous namespace).
#pragma once

constexpr int kLossThreads = 256;
// layout of the output / saved-for-backward record
enum {
    PL_PLOSS = 0, PL_WLOSS, PL_ALOSS, PL_RLOSS, PL_S_WEIGHT, PL_Q_WEIGHT, PL_S_ACT, PL_Q_ACT,
    PL_WEIGHT_REG, PL_WACT, PL_AACT, PL_CW, PL_CA, PL_L2, PL_NOUT
};

struct PLossArgs {
    const float *las, *laq;     // [n_a] log_act_s, log_act_q
    const float *lws, *lwq;     // [n_w] log_wght_s, log2(row range + 2^log_wght_s)
    int64_t n_a, n_w;
    float wt, at, eps, t;       // target bit widths, l_eps, temperature
    int lossless, training;
};

__device__ __forceinline__ double block_sum_f64(double v, double *s) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    v = warp_sum_f64(v);
    if (lane == 0) s[w] = v;
    __syncthreads();
    double r = 0.0;
    for (int i = 0; i < kLossThreads / 32; ++i) r += s[i];     // fixed order, every thread
    __syncthreads();
    return r;
}
__device__ __forceinline__ float block_max_f32(float v, float *s) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    v = warp_max_nan(v);
    if (lane == 0) s[w] = v;
    __syncthreads();
    float r = s[0];
    for (int i = 1; i < kLossThreads / 32; ++i) r = max_nan(r, s[i]);
    __syncthreads();
    return r;
}

// One side of the constraint: mean of max(0, (q - s) - thr), the active count, and the means of
// s and q themselves (the logged s_*_loss / q_*_loss terms).
__device__ __forceinline__ void pl_side(const float *ls, const float *lq, int64_t n, float thr, double *s64,
                                        float *s32, float &loss, float &act, float &mean_s, float &mean_q,
                                        float &max_d) {
    double sl = 0.0, sa = 0.0, ss = 0.0, sq = 0.0;
    float md = -INFINITY;
    for (int64_t i = threadIdx.x; i < n; i += kLossThreads) {
        const float s = ls[i], q = lq[i];
        const float d = __fsub_rn(q, s);                 // lwq - lws
        const float x = __fsub_rn(d, thr);
        const float l0 = x > 0.f ? x : (x == x ? 0.f : x);   // torch.max(0, x): NaN propagates
        sl += (double)l0;
        sa += (l0 > 0.f) ? 1.0 : 0.0;
        ss += (double)s;
        sq += (double)q;
        md = max_nan(md, d);
    }
    sl = block_sum_f64(sl, s64);
    sa = block_sum_f64(sa, s64);
    ss = block_sum_f64(ss, s64);
    sq = block_sum_f64(sq, s64);
    max_d = block_max_f32(md, s32);
    const double inv = n > 0 ? 1.0 / (double)n : 0.0 / 0.0;
    loss = (float)(sl * inv);
    act = (float)sa;
    mean_s = (float)(ss * inv);
    mean_q = (float)(sq * inv);
}

__global__ void __launch_bounds__(kLossThreads)
fq_potential_loss_fwd_kernel(PLossArgs a, const float *__restrict__ base_loss, float *loss_sum, float *cnt,
                             float *__restrict__ out) {
    __shared__ double s64[kLossThreads / 32];
    __shared__ float s32[kLossThreads / 32];
    const float thr_w = __fsub_rn(a.wt, a.eps), thr_a = __fsub_rn(a.at, a.eps);
    float wloss, wact, ms_w, mq_w, maxd_w, aloss, aact, ms_a, mq_a, maxd_a;
    pl_side(a.lws, a.lwq, a.n_w, thr_w, s64, s32, wloss, wact, ms_w, mq_w, maxd_w);
    pl_side(a.las, a.laq, a.n_a, thr_a, s64, s32, aloss, aact, ms_a, mq_a, maxd_a);
    if (threadIdx.x != 0) return;
    const float rloss = base_loss[0];                         // pow_(p = 1)
    const float calib = __fdiv_rn(loss_sum[0], cnt[0]);
    const float den = __fadd_rn(__fadd_rn(wact, aact), a.eps);
    const float wmul = __fdiv_rn(__fadd_rn(wact, a.eps), den);
    const float amul = __fdiv_rn(__fadd_rn(aact, a.eps), den);
    const float l1 = a.lossless ? 1.f : a.t, l2 = a.lossless ? a.t : 1.f;
    const float cl = __fmul_rn(calib, l1);
    const float mix = __fadd_rn(__fmul_rn(wmul, wloss), __fmul_rn(amul, aloss));
    out[PL_PLOSS] = __fadd_rn(__fmul_rn(cl, mix), __fmul_rn(l2, rloss));
    out[PL_WLOSS] = wloss;
    out[PL_ALOSS] = aloss;
    out[PL_RLOSS] = rloss;
    out[PL_S_WEIGHT] = -ms_w;
    out[PL_Q_WEIGHT] = mq_w;
    out[PL_S_ACT] = -ms_a;
    out[PL_Q_ACT] = mq_a;
    out[PL_WEIGHT_REG] = maxd_w;
    out[PL_WACT] = wact;
    out[PL_AACT] = aact;
    // d ploss / d wloss0_i and d ploss / d aloss0_j (the means' 1/n folded in), for the backward
    out[PL_CW] = a.n_w > 0 ? __fdiv_rn(__fmul_rn(cl, wmul), (float)a.n_w) : 0.f;
    out[PL_CA] = a.n_a > 0 ? __fdiv_rn(__fmul_rn(cl, amul), (float)a.n_a) : 0.f;
    out[PL_L2] = l2;
    if (a.training) {                                         // gdnsq_loss.py:73-75
        loss_sum[0] = __fadd_rn(loss_sum[0], rloss);
        cnt[0] = __fadd_rn(cnt[0], 1.f);
    }
}

// d ploss / d {las, laq, lws, lwq, base_loss} times the upstream gradient g[0].
// torch.max(z, x) passes the gradient to x where x > z and HALF of it on a tie (x == z).
__global__ void __launch_bounds__(kLossThreads)
fq_potential_loss_bwd_kernel(PLossArgs a, const float *__restrict__ saved, const float *__restrict__ g,
                             float *__restrict__ g_las, float *__restrict__ g_laq,
                             float *__restrict__ g_lws, float *__restrict__ g_lwq,
                             float *__restrict__ g_base) {
    const float go = g[0];
    const float cw = __fmul_rn(go, saved[PL_CW]), ca = __fmul_rn(go, saved[PL_CA]);
    const float thr_w = __fsub_rn(a.wt, a.eps), thr_a = __fsub_rn(a.at, a.eps);
    const int64_t i0 = (int64_t)blockIdx.x * kLossThreads + threadIdx.x, stride = (int64_t)gridDim.x * kLossThreads;
    for (int64_t i = i0; i < a.n_w; i += stride) {
        const float x = __fsub_rn(__fsub_rn(a.lwq[i], a.lws[i]), thr_w);
        const float m = x > 0.f ? 1.f : (x == 0.f ? 0.5f : 0.f);
        const float v = __fmul_rn(cw, m);
        if (g_lwq) g_lwq[i] = v;
        if (g_lws) g_lws[i] = -v;
    }
    for (int64_t i = i0; i < a.n_a; i += stride) {
        const float x = __fsub_rn(__fsub_rn(a.laq[i], a.las[i]), thr_a);
        const float m = x > 0.f ? 1.f : (x == 0.f ? 0.5f : 0.f);
        const float v = __fmul_rn(ca, m);
        if (g_laq) g_laq[i] = v;
        if (g_las) g_las[i] = -v;
    }
    if (g_base && i0 == 0) g_base[0] = __fmul_rn(go, saved[PL_L2]);
}
