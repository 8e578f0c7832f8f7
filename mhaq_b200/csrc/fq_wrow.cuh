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

template <bool VEC>
__global__ void __launch_bounds__(kThreads)
fq_wrow_fwd_kernel(const float *__restrict__ w, float *__restrict__ wq, WRowArgs a,
                   float *__restrict__ row_min, float *__restrict__ row_max,
                   float *__restrict__ log_range) {
    __shared__ RowStat s_w[kThreads / 32];
    __shared__ float s_b[2];
    const int tid = threadIdx.x;
    for (int64_t row = blockIdx.x; row < a.n_rows; row += gridDim.x) {
        const float *xr = w + row * a.n_inner;
        // ---- pass 1: row minimum / maximum (same merge as fq_rowstat_kernel) ----
        RowStat st = {INFINITY, -INFINITY, 0.f, 0.f};
#pragma unroll 4
        for (int64_t p = (int64_t)tid * 4; p < a.n_inner; p += kIterElems) {
            const int nv = valid4<VEC>(p, a.n_inner);
            const float4 v = load4<VEC>(xr, p, a.n_inner);
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
        const float s = exp2f(__ldg(a.log_s + row * a.ls));
        if (tid == 0) {
            RowStat z = s_w[0];
            for (int i = 1; i < kThreads / 32; ++i) rs_merge(z, s_w[i]);
            s_b[0] = z.mn;
            s_b[1] = z.mx;
            if (row_min) row_min[row] = z.mn;
            if (row_max) row_max[row] = z.mx;
            // torch.log2(mx - mn + torch.exp2(log_wght_s)): three fp32 roundings, then log2
            if (log_range) log_range[row] = log2f(f_add(f_sub(z.mx, z.mn), s));
        }
        __syncthreads();
        // ---- pass 2: quantize the row (second read comes from L2) ----
        if (wq) {
            QConst q;
            q.s = s;
            q.zp = s_b[0];
            q.lo = -INFINITY;
            q.hi = INFINITY;
            float *yr = wq + row * a.n_inner;
#pragma unroll 4
            for (int64_t p = (int64_t)tid * 4; p < a.n_inner; p += kIterElems) {
                const float4 v = load4<VEC>(xr, p, a.n_inner);
                float4 y;
                float code;
                y.x = fwd_elem(v.x, q, code);
                y.y = fwd_elem(v.y, q, code);
                y.z = fwd_elem(v.z, q, code);
                y.w = fwd_elem(v.w, q, code);
                store4<VEC>(yr, p, a.n_inner, y);
            }
        }
        __syncthreads();      // s_w / s_b are reused by the next row
    }
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
    const int tid = threadIdx.x;
    PhiloxKey key = {0, 0, 0, 0};
    if (NOISE == NOISE_PHILOX) key = make_key(seed, offset, philox_dev);
    const int64_t supers_per_row = (a.n_inner + kSuperElems - 1) / kSuperElems;
    const int64_t subs_per_row = (a.n_inner + kSubElems - 1) / kSubElems;
    constexpr float kLn2f = 0.6931471805599453f;
    constexpr double kLn2 = 0.693147180559945309417;
    const bool has_max = (g_log_range != nullptr) || (g_row_max != nullptr);

    for (int64_t row = blockIdx.x; row < a.n_rows; row += gridDim.x) {
        const int64_t off = row * a.n_inner;
        const float *xr = w + off, *gr = go + off;
        const float *rr = (NOISE == NOISE_EXPLICIT) ? r + off : nullptr;
        float *or_ = gw ? gw + off : nullptr;
        const float mn = __ldg(row_min + row), mx = __ldg(row_max + row);
        QConst q;
        q.s = exp2f(__ldg(a.log_s + row * a.ls));
        q.zp = mn;
        q.lo = -INFINITY;
        q.hi = INFINITY;
        BwdConst bc;
        bc.smul = q.s;
        bc.rcp = __frcp_rn(q.s);
        bc.delta = 0.f;
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
                    rnd = noise_block(key, row, supers_per_row, T, tid);
                    curT = T;
                }
            }
            const int it0 = (int)(sub & (kSuperSubs - 1)) * kSubIters;
#pragma unroll 4
            for (int it = 0; it < kSubIters; ++it) {
                const int64_t p = sub * kSubElems + (int64_t)it * kIterElems + tid * 4;
                const int nv = valid4<VEC>(p, a.n_inner);
                if (!nv) continue;
                const float4 xv = load4<VEC>(xr, p, a.n_inner);
                const float4 gv = load4<VEC>(gr, p, a.n_inner);
                float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (NOISE == NOISE_EXPLICIT) rv = load4<VEC>(rr, p, a.n_inner);
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
                if (or_) store4<VEC>(or_, p, a.n_inner, o);
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
            if (g_log_range) {
                // autograd of log2((mx - mn) + exp2(log_s)) in torch's fp32 op order:
                //   Log2Backward  grad / (self * ln2);  Exp2Backward  grad * result * ln2
                const float tt = f_add(f_sub(mx, mn), q.s);
                gt = f_div(__ldg(g_log_range + row), f_mul(tt, kLn2f));
                g2 = f_mul(f_mul(gt, q.s), kLn2f);
            }
            if (g_log_s) g_log_s[row] = g_log_range ? f_add(g1, g2) : g1;
            // amin / amax backward: what flows into the row minimum and maximum
            float gmin = dzp;                                    // zero point = row minimum
            if (g_log_range) gmin = f_add(gmin, -gt);            // SubBackward of (mx - mn)
            if (g_row_min) gmin = f_add(gmin, __ldg(g_row_min + row));
            float gmax = g_log_range ? gt : 0.f;
            if (g_row_max) gmax = g_log_range ? f_add(gmax, __ldg(g_row_max + row)) : __ldg(g_row_max + row);
            s_d[0] = f_div(gmin, (float)t[3]);                   // even split among ties
            s_d[1] = has_max ? f_div(gmax, (float)t[4]) : 0.f;
        }
        __syncthreads();
        // ---- pass 2: patch the tie elements (each thread re-visits the float4s it wrote) ----
        if (or_) {
            const float dmn = s_d[0], dmx = s_d[1];
            for (int64_t p = (int64_t)tid * 4; p < a.n_inner; p += kIterElems) {
                const int nv = valid4<VEC>(p, a.n_inner);
                const float4 xv = load4<VEC>(xr, p, a.n_inner);
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
}
