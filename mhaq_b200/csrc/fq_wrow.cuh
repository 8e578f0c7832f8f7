// fq_wrow.cuh — row-resident fused kernels for per-channel WEIGHT quantizers.
// (Included inside the anonymous namespace of mhaq_fq.cu: uses fwd_elem / bwd_elem / RowStat.)
//
// A conv / linear weight is [O rows][I*kh*kw] with a few hundred to a few thousand elements per
// row and one quantization channel per row.  The streaming kernels treat it like any tensor:
// row statistics, forward, backward, finalize and the amin/amax scatter are five launches, and
// ModelHelper.get_model_values (utils/model_helper.py:24-25,44) adds ~10 more tiny elementwise
// launches per layer and step for log2(max - min + 2^log_wght_s) and its autograd.  For such
// short rows everything a channel needs is inside ONE CTA, so the whole weight path is
//
//   fq_wrow_fwd_kernel : row min / max  ->  zp = min, s = exp2(log_wght_s[row])
//                        wq = fake_quant(w; s, zp)               (gdnsq_conv2d.py:72-84, 98)
//                        log_range[row] = log2((max - min) + s)  (model_helper.py:24-25,44)
//   fq_wrow_bwd_kernel : gx and the channel's parameter sums in one pass (same per-element
//                        arithmetic as fq_bwd_kernel's general path), reduced inside the CTA
//                        (fp32 per thread -> warp shuffle -> fp64 across warps, fixed order),
//                        d/d log_wght_s (both uses: the quantizer and log_range), then the
//                        amin / amax backward — even split among ties — patched onto the row.
//
// One launch each way instead of ~6 and ~11.  Same element -> thread mapping as the streaming
// kernels (thread tid owns the float4 at it*512 + tid*4), so the in-kernel noise stream is the
// same function of (seed, offset, row, position).

struct WRowArgs {
    const float *log_s;
    int ls;                 // stride of log_s: 0 or 1
    int64_t n_rows, n_inner;
};

// One row of the forward (one CTA): row_min / row_max / log_range point at THIS row's outputs.
template <bool VEC>
__device__ __forceinline__ void wrow_fwd_row(const float *__restrict__ w_row, float *__restrict__ wq,
                                             float log_s, int64_t n_inner, float *row_min, float *row_max,
                                             float *log_range, RowStat *s_w, float *s_b) {
    const int tid = threadIdx.x;
    const float *xr = w_row;
    // ---- pass 1: row minimum / maximum (same merge as fq_rowstat_kernel) ----
    RowStat st = {INFINITY, -INFINITY, 0.f, 0.f};
#pragma unroll 4
    for (int64_t p = (int64_t)tid * 4; p < n_inner; p += kIterElems) {
        const int nv = valid4<VEC>(p, n_inner);
        const float4 v = load4<VEC>(xr, p, n_inner);
        rs_push(st, v.x);
        if (nv > 1) rs_push(st, v.y);
        if (nv > 2) rs_push(st, v.z);
        if (nv > 3) rs_push(st, v.w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        RowStat b;
        b.mn = __shfl_xor_sync(0xffffffffu, st.mn, o);
        b.mx = __shfl_xor_sync(0xffffffffu, st.mx, o);
        b.cmn = __shfl_xor_sync(0xffffffffu, st.cmn, o);
        b.cmx = __shfl_xor_sync(0xffffffffu, st.cmx, o);
        rs_merge(st, b);
    }
    if ((tid & 31) == 0) s_w[tid >> 5] = st;
    __syncthreads();
    const float s = exp2f(log_s);
    if (tid == 0) {
        RowStat z = s_w[0];
        for (int i = 1; i < kThreads / 32; ++i) rs_merge(z, s_w[i]);
        s_b[0] = z.mn;
        s_b[1] = z.mx;
        if (row_min) *row_min = z.mn;
        if (row_max) *row_max = z.mx;
        // torch.log2(mx - mn + torch.exp2(log_wght_s)): three fp32 roundings, then log2
        if (log_range) *log_range = log2f(f_add(f_sub(z.mx, z.mn), s));
    }
    __syncthreads();
    // ---- pass 2: quantize the row (second read comes from L2) ----
    if (wq) {
        QConst q;
        q.s = s;
        q.zp = s_b[0];
        q.lo = -INFINITY;
        q.hi = INFINITY;
        float *yr = wq;
#pragma unroll 4
        for (int64_t p = (int64_t)tid * 4; p < n_inner; p += kIterElems) {
            const float4 v = load4<VEC>(xr, p, n_inner);
            float4 y;
            float code;
            y.x = fwd_elem(v.x, q, code);
            y.y = fwd_elem(v.y, q, code);
            y.z = fwd_elem(v.z, q, code);
            y.w = fwd_elem(v.w, q, code);
            store4<VEC>(yr, p, n_inner, y);
        }
    }
    __syncthreads();      // s_w / s_b are reused by the next row
}

template <bool VEC>
__global__ void __launch_bounds__(kThreads)
fq_wrow_fwd_kernel(const float *__restrict__ w, float *__restrict__ wq, WRowArgs a,
                   float *__restrict__ row_min, float *__restrict__ row_max,
                   float *__restrict__ log_range) {
    __shared__ RowStat s_w[kThreads / 32];
    __shared__ float s_b[2];
    for (int64_t row = blockIdx.x; row < a.n_rows; row += gridDim.x)
        wrow_fwd_row<VEC>(w + row * a.n_inner, wq ? wq + row * a.n_inner : nullptr,
                          __ldg(a.log_s + row * a.ls), a.n_inner, row_min ? row_min + row : nullptr,
                          row_max ? row_max + row : nullptr, log_range ? log_range + row : nullptr, s_w, s_b);
}

// ---- multi-tensor launch: every per-channel weight of a model in ONE grid ----------------
// The descriptors travel in the kernel parameter space (<= kWRowMultiMax tensors per launch, 4 KB
// parameter limit); block b finds its (tensor, row) from the prefix sums of the row counts.
constexpr int kWRowMultiMax = 32;
struct WRowFwdDesc {
    const float *w, *log_s;
    float *wq, *row_min, *row_max, *log_range;
    int64_t n_inner;
};
struct WRowFwdMulti {
    WRowFwdDesc d[kWRowMultiMax];
    int row0[kWRowMultiMax + 1];      // prefix sums of the row counts
    int vec[kWRowMultiMax];           // 128-bit path allowed (alignment, n_inner % 4 == 0)
    int n;
};
__device__ __forceinline__ int wrow_find(const int *row0, int n, int row) {
    int t = 0;
    while (t + 1 < n && row >= row0[t + 1]) ++t;
    return t;
}

__global__ void __launch_bounds__(kThreads)
fq_wrow_multi_fwd_kernel(const __grid_constant__ WRowFwdMulti m) {
    __shared__ RowStat s_w[kThreads / 32];
    __shared__ float s_b[2];
    const int total = m.row0[m.n];
    for (int row = blockIdx.x; row < total; row += gridDim.x) {
        const int t = wrow_find(m.row0, m.n, row);
        const WRowFwdDesc &d = m.d[t];
        const int64_t r = row - m.row0[t];
        const float ls = __ldg(d.log_s + r);
        const float *w_row = d.w + r * d.n_inner;
        float *wq = d.wq ? d.wq + r * d.n_inner : nullptr;
        float *mn = d.row_min ? d.row_min + r : nullptr, *mx = d.row_max ? d.row_max + r : nullptr;
        float *lr = d.log_range ? d.log_range + r : nullptr;
        if (m.vec[t]) wrow_fwd_row<true>(w_row, wq, ls, d.n_inner, mn, mx, lr, s_w, s_b);
        else wrow_fwd_row<false>(w_row, wq, ls, d.n_inner, mn, mx, lr, s_w, s_b);
    }
}

// Everything one row of the backward needs, resolved by the caller (single- or multi-tensor).
struct WRowBwdRow {
    const float *go, *w, *r;     // this row's upstream gradient, weights, explicit noise (or NULL)
    float *gw, *g_log_s;         // this row's outputs (either may be NULL)
    float log_s, mn, mx;         // log_wght_s[row], row minimum / maximum from the forward
    float g_lr, g_mn, g_mx;      // gradients w.r.t. log_range / row_min / row_max of this row
    float g_ls_acc;              // gradient log_wght_s[row] receives along the caller's other paths
    bool has_acc;
    float delta;                 // AEWGS: this channel's delta from the (all-reduced) statistics
    bool has_glr, has_gmn, has_gmx;
    int64_t noise_row;           // row index inside ITS tensor: the noise stream's coordinate
};

template <int METHOD, int NOISE, bool VEC>
__device__ __forceinline__ void wrow_bwd_row(const WRowBwdRow &d, int64_t n_inner, const PhiloxKey &key,
                                             float (*s_red)[kThreads / 32], float *s_d) {
    const int tid = threadIdx.x;
    const int64_t supers_per_row = (n_inner + kSuperElems - 1) / kSuperElems;
    const int64_t subs_per_row = (n_inner + kSubElems - 1) / kSubElems;
    constexpr float kLn2f = 0.6931471805599453f;
    constexpr double kLn2 = 0.693147180559945309417;
    const bool has_max = d.has_glr || d.has_gmx;
    const float *xr = d.w, *gr = d.go;
    const float *rr = (NOISE == NOISE_EXPLICIT) ? d.r : nullptr;
    float *or_ = d.gw;
    const float mn = d.mn, mx = d.mx;
    QConst q;
    q.s = exp2f(d.log_s);
    q.zp = mn;
    q.lo = -INFINITY;
    q.hi = INFINITY;
    BwdConst bc;
    bc.smul = q.s;
    bc.rcp = __frcp_rn(q.s);
    bc.delta = d.delta;
    bc.lo_lt_hi = true;
    bc.lo_gt_hi = false;
    Acc acc = {0.f, 0.f, 0.f, 0.f, 0.f};
    float cmn = 0.f, cmx = 0.f;
    uint4 rnd = make_uint4(0, 0, 0, 0);
    int64_t curT = -1;
    // ---- pass 1: input gradient + the channel's partial sums + tie counts ----
    for (int64_t sub = 0; sub < subs_per_row; ++sub) {
        if (NOISE == NOISE_PHILOX) {
            const int64_t T = sub / kSuperSubs;
            if (T != curT) {
                rnd = noise_block(key, d.noise_row, supers_per_row, T, tid);
                curT = T;
            }
        }
        const int it0 = (int)(sub & (kSuperSubs - 1)) * kSubIters;
#pragma unroll 4
        for (int it = 0; it < kSubIters; ++it) {
            const int64_t p = sub * kSubElems + (int64_t)it * kIterElems + tid * 4;
            const int nv = valid4<VEC>(p, n_inner);
            if (!nv) continue;
            const float4 xv = load4<VEC>(xr, p, n_inner);
            const float4 gv = load4<VEC>(gr, p, n_inner);
            float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (NOISE == NOISE_EXPLICIT) rv = load4<VEC>(rr, p, n_inner);
            uint32_t inv = 0;
            if (NOISE == NOISE_PHILOX) inv = ~noise_nibble(rnd, it0 + it);
            float4 o;
            o.x = bwd_elem<METHOD, false, NOISE, false, false>(xv.x, gv.x, rv.x, inv << 31, q, bc, acc);
            o.y = bwd_elem<METHOD, false, NOISE, false, false>(xv.y, gv.y, rv.y, inv << 30, q, bc, acc);
            o.z = bwd_elem<METHOD, false, NOISE, false, false>(xv.z, gv.z, rv.z, inv << 29, q, bc, acc);
            o.w = bwd_elem<METHOD, false, NOISE, false, false>(xv.w, gv.w, rv.w, inv << 28, q, bc, acc);
            const float xe[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (e < nv) {
                    cmn += (xe[e] == mn) ? 1.f : 0.f;
                    cmx += (xe[e] == mx) ? 1.f : 0.f;
                }
            }
            if (or_) store4<VEC>(or_, p, n_inner, o);
        }
    }
    // ---- channel sums: fp32 per thread -> warp shuffle -> fp64 over the warps ----
    float v0 = warp_sum(acc.se), v1 = warp_sum(acc.sn), v2 = warp_sum(acc.sz);
    float v3 = warp_sum(cmn), v4 = warp_sum(cmx);
    if ((tid & 31) == 0) {
        const int wi = tid >> 5;
        s_red[0][wi] = v0; s_red[1][wi] = v1; s_red[2][wi] = v2; s_red[3][wi] = v3; s_red[4][wi] = v4;
    }
    __syncthreads();
    if (tid == 0) {
        double t[5];
#pragma unroll
        for (int m = 0; m < 5; ++m) {
            t[m] = 0.0;
#pragma unroll
            for (int wi = 0; wi < kThreads / 32; ++wi) t[m] += (double)s_red[m][wi];
        }
        // d/d log_wght_s through the quantizer: (S_e + S_noise) * s * ln2   (Exp2Backward)
        const float g1 = (float)((t[0] + t[1]) * (double)q.s * kLn2);
        const float dzp = (float)t[2];              // d/d zero_point = sum(go - g_u)
        float g2 = 0.f, gt = 0.f;
        if (d.has_glr) {
            // autograd of log2((mx - mn) + exp2(log_s)) in torch's fp32 op order:
            //   Log2Backward  grad / (self * ln2);  Exp2Backward  grad * result * ln2
            const float tt = f_add(f_sub(mx, mn), q.s);
            gt = f_div(d.g_lr, f_mul(tt, kLn2f));
            g2 = f_mul(f_mul(gt, q.s), kLn2f);
        }
        if (d.g_log_s) {
            float gls = d.has_glr ? f_add(g1, g2) : g1;
            if (d.has_acc) gls = f_add(gls, d.g_ls_acc);    // == autograd's AccumulateGrad of the two
            *d.g_log_s = gls;
        }
        // amin / amax backward: what flows into the row minimum and maximum
        float gmin = dzp;                                    // zero point = row minimum
        if (d.has_glr) gmin = f_add(gmin, -gt);            // SubBackward of (mx - mn)
        if (d.has_gmn) gmin = f_add(gmin, d.g_mn);
        float gmax = d.has_glr ? gt : 0.f;
        if (d.has_gmx) gmax = d.has_glr ? f_add(gmax, d.g_mx) : d.g_mx;
        s_d[0] = f_div(gmin, (float)t[3]);                   // even split among ties
        s_d[1] = has_max ? f_div(gmax, (float)t[4]) : 0.f;
    }
    __syncthreads();
    // ---- pass 2: patch the tie elements (each thread re-visits the float4s it wrote) ----
    if (or_) {
        const float dmn = s_d[0], dmx = s_d[1];
        for (int64_t p = (int64_t)tid * 4; p < n_inner; p += kIterElems) {
            const int nv = valid4<VEC>(p, n_inner);
            const float4 xv = load4<VEC>(xr, p, n_inner);
            const float xe[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (e < nv) {
                    const bool at_mn = xe[e] == mn, at_mx = has_max && (xe[e] == mx);
                    if (at_mn || at_mx) {
                        float o = or_[p + e];
                        if (at_mn) o = f_add(o, dmn);
                        if (at_mx) o = f_add(o, dmx);
                        or_[p + e] = o;
                    }
                }
            }
        }
    }
    __syncthreads();      // s_red / s_d are reused by the next row
}

template <int METHOD, int NOISE, bool VEC>
__global__ void __launch_bounds__(kThreads)
fq_wrow_bwd_kernel(const float *__restrict__ go, const float *__restrict__ w, WRowArgs a,
                   const float *__restrict__ row_min, const float *__restrict__ row_max,
                   const float *__restrict__ g_log_range, const float *__restrict__ g_row_min,
                   const float *__restrict__ g_row_max, const float *__restrict__ r, uint64_t seed,
                   uint64_t offset, const uint64_t *__restrict__ philox_dev, float *gw,
                   float *__restrict__ g_log_s) {
    __shared__ float s_red[5][kThreads / 32];
    __shared__ float s_d[2];
    PhiloxKey key = {0, 0, 0, 0};
    if (NOISE == NOISE_PHILOX) key = make_key(seed, offset, philox_dev);
    for (int64_t row = blockIdx.x; row < a.n_rows; row += gridDim.x) {
        const int64_t off = row * a.n_inner;
        WRowBwdRow d;
        d.go = go + off; d.w = w + off;
        d.r = (NOISE == NOISE_EXPLICIT) ? r + off : nullptr;
        d.gw = gw ? gw + off : nullptr;
        d.g_log_s = g_log_s ? g_log_s + row : nullptr;
        d.log_s = __ldg(a.log_s + row * a.ls);
        d.mn = __ldg(row_min + row); d.mx = __ldg(row_max + row);
        d.has_glr = g_log_range != nullptr; d.has_gmn = g_row_min != nullptr; d.has_gmx = g_row_max != nullptr;
        d.g_lr = d.has_glr ? __ldg(g_log_range + row) : 0.f;
        d.g_mn = d.has_gmn ? __ldg(g_row_min + row) : 0.f;
        d.g_mx = d.has_gmx ? __ldg(g_row_max + row) : 0.f;
        d.noise_row = row;
        d.delta = 0.f;
        d.has_acc = false; d.g_ls_acc = 0.f;
        wrow_bwd_row<METHOD, NOISE, VEC>(d, a.n_inner, key, s_red, s_d);
    }
}

// ---- multi-tensor backward: every per-channel weight of a model in ONE grid ---------------
struct WRowBwdDesc {
    const float *g_wq, *w, *log_s, *row_min, *row_max, *g_log_range, *g_row_min, *g_row_max, *r;
    float *g_w, *g_log_s;
    int64_t n_inner;
    const float *g_log_s_acc;
};
constexpr int kWRowMultiBwdMax = 24;      // 24 x 104 B of descriptors + tables < 4 KB of parameters
struct WRowBwdMulti {
    WRowBwdDesc d[kWRowMultiBwdMax];
    int row0[kWRowMultiBwdMax + 1];
    int vec[kWRowMultiBwdMax];
    int n;
};

// Tensor t draws its noise from Philox stream (seed, offset + t): the stream a per-layer launch
// with that offset would read.
// AEWGS: `stats` = packed per-row means [3][total rows of the launch's model] (num | e2 | me), already
// averaged over the ranks; `stats_row0` = index of this launch's first row in it, `stats_ld` = its
// leading dimension (gdnsq.py:126-134).
template <int METHOD, int NOISE>
__global__ void __launch_bounds__(kThreads)
fq_wrow_multi_bwd_kernel(const __grid_constant__ WRowBwdMulti m, uint64_t seed, uint64_t offset,
                         const uint64_t *__restrict__ philox_dev, const float *__restrict__ stats,
                         int64_t stats_row0, int64_t stats_ld) {
    __shared__ float s_red[5][kThreads / 32];
    __shared__ float s_d[2];
    const int total = m.row0[m.n];
    for (int row = blockIdx.x; row < total; row += gridDim.x) {
        const int t = wrow_find(m.row0, m.n, row);
        const WRowBwdDesc &e = m.d[t];
        const int64_t r = row - m.row0[t];
        const int64_t off = r * e.n_inner;
        PhiloxKey key = {0, 0, 0, 0};
        if (NOISE == NOISE_PHILOX) key = make_key(seed, offset + (uint64_t)t, philox_dev);
        WRowBwdRow d;
        d.go = e.g_wq + off; d.w = e.w + off;
        d.r = (NOISE == NOISE_EXPLICIT) ? e.r + off : nullptr;
        d.gw = e.g_w ? e.g_w + off : nullptr;
        d.g_log_s = e.g_log_s ? e.g_log_s + r : nullptr;
        d.log_s = __ldg(e.log_s + r);
        d.mn = __ldg(e.row_min + r); d.mx = __ldg(e.row_max + r);
        d.has_glr = e.g_log_range != nullptr; d.has_gmn = e.g_row_min != nullptr; d.has_gmx = e.g_row_max != nullptr;
        d.g_lr = d.has_glr ? __ldg(e.g_log_range + r) : 0.f;
        d.g_mn = d.has_gmn ? __ldg(e.g_row_min + r) : 0.f;
        d.g_mx = d.has_gmx ? __ldg(e.g_row_max + r) : 0.f;
        d.noise_row = r;
        d.delta = 0.f;
        d.has_acc = e.g_log_s_acc != nullptr;
        d.g_ls_acc = d.has_acc ? __ldg(e.g_log_s_acc + r) : 0.f;
        if (METHOD == MHAQ_FQ_AEWGS) {
            const int64_t sr = stats_row0 + row;
            const float num = __ldg(stats + sr), e2 = __ldg(stats + stats_ld + sr), me = __ldg(stats + 2 * stats_ld + sr);
            float den = f_sub(e2, f_mul(me, me));                 // gdnsq.py:132
            den = (den < kAewgsEps) ? kAewgsEps : den;            // clamp_min(eps)
            d.delta = f_div(num, den);                            // gdnsq.py:134
        }
        if (m.vec[t]) wrow_bwd_row<METHOD, NOISE, true>(d, e.n_inner, key, s_red, s_d);
        else wrow_bwd_row<METHOD, NOISE, false>(d, e.n_inner, key, s_red, s_d);
    }
}

// ---- AEWGS statistics of every row of every tensor, one grid (gdnsq.py:118-124) --------------
// stats[0][row] = mean(sign(g) * e), stats[1][row] = mean(e * e), stats[2][row] = mean(e) over the
// row, g = go * s, e = round(v) - v: the packed buffer the caller all-reduces ONCE for the whole
// model (the reference: three all-reduces per weight tensor, gdnsq.py:126-129).
template <bool VEC>
__device__ __forceinline__ void wrow_aewgs_stats_row(const float *__restrict__ w_row,
                                                     const float *__restrict__ go_row, float log_s, float mn,
                                                     int64_t n_inner, float *o_num, float *o_e2, float *o_me,
                                                     float (*s_red)[kThreads / 32]) {
    const int tid = threadIdx.x;
    const float s = exp2f(log_s);
    float a_num = 0.f, a_e2 = 0.f, a_e = 0.f;
    for (int64_t p = (int64_t)tid * 4; p < n_inner; p += kIterElems) {
        const int nv = valid4<VEC>(p, n_inner);
        const float4 xv = load4<VEC>(w_row, p, n_inner);
        const float4 gv = load4<VEC>(go_row, p, n_inner);
        const float xe[4] = {xv.x, xv.y, xv.z, xv.w}, ge[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (i < nv) {
                const float v = f_div(f_sub(xe[i], mn), s);
                const float e = f_sub(rintf(v), v);
                const float gg = f_mul(ge[i], s);
                const float sg = (gg > 0.f) ? 1.f : ((gg < 0.f) ? -1.f : 0.f);
                a_num += f_mul(sg, e);
                a_e2 += f_mul(e, e);
                a_e += e;
            }
        }
    }
    const float v0 = warp_sum(a_num), v1 = warp_sum(a_e2), v2 = warp_sum(a_e);
    if ((tid & 31) == 0) {
        const int wi = tid >> 5;
        s_red[0][wi] = v0; s_red[1][wi] = v1; s_red[2][wi] = v2;
    }
    __syncthreads();
    if (tid == 0) {
        double t[3];
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            t[m] = 0.0;
#pragma unroll
            for (int wi = 0; wi < kThreads / 32; ++wi) t[m] += (double)s_red[m][wi];
        }
        const double cnt = (double)n_inner;
        *o_num = (float)(t[0] / cnt);
        *o_e2 = (float)(t[1] / cnt);
        *o_me = (float)(t[2] / cnt);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kThreads)
fq_wrow_multi_aewgs_stats_kernel(const __grid_constant__ WRowBwdMulti m, float *__restrict__ stats,
                                 int64_t stats_row0, int64_t stats_ld) {
    __shared__ float s_red[5][kThreads / 32];
    const int total = m.row0[m.n];
    for (int row = blockIdx.x; row < total; row += gridDim.x) {
        const int t = wrow_find(m.row0, m.n, row);
        const WRowBwdDesc &e = m.d[t];
        const int64_t r = row - m.row0[t];
        const int64_t off = r * e.n_inner;
        const int64_t sr = stats_row0 + row;
        float *o0 = stats + sr, *o1 = stats + stats_ld + sr, *o2 = stats + 2 * stats_ld + sr;
        if (m.vec[t])
            wrow_aewgs_stats_row<true>(e.w + off, e.g_wq + off, __ldg(e.log_s + r), __ldg(e.row_min + r), e.n_inner,
                                       o0, o1, o2, s_red);
        else
            wrow_aewgs_stats_row<false>(e.w + off, e.g_wq + off, __ldg(e.log_s + r), __ldg(e.row_min + r), e.n_inner,
                                        o0, o1, o2, s_red);
    }
}
