// fq_common.cuh — shared device helpers for the sm_100a fake-quant kernels.
//
// Work decomposition (shared by every streaming kernel in this library)
// ---------------------------------------------------------------------
// A tensor is [n_rows][n_inner] fp32, row r quantized with channel r % n_ch.
// Each row is cut into
//     sub-tiles   of 4096 elements  (= 128 threads x 8 iterations x float4)
//     super-tiles of 16384 elements (= 4 sub-tiles = 32 iterations)
// Inside a super-tile, thread `tid` of the 128-thread CTA owns the float4 at
// element offset  it*512 + tid*4  for it = 0..31.  Every warp-level access is
// therefore 512 contiguous bytes (four full 128-byte lines).
// A *task* is `spt` consecutive sub-tiles of one row; one CTA runs one task at
// a time and, in the reducing kernels, flushes one partial record per task, so
// the reduction tree is a pure function of the tensor shape (deterministic,
// no atomics).  The noise stream is defined on (row, position-in-row): one
// Philox4x32-10 block (128 bits) covers the 32 float4s a thread owns in a
// super-tile, bit 4*it+k for element k of iteration it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace mhaq {

constexpr int kThreads = 128;              // threads per CTA
constexpr int kIterElems = kThreads * 4;   // 512 elements per CTA iteration
constexpr int kSubIters = 8;               // iterations per sub-tile
constexpr int kSubElems = kIterElems * kSubIters;  // 4096
constexpr int kSuperSubs = 4;              // sub-tiles per super-tile
constexpr int kSuperElems = kSubElems * kSuperSubs;  // 16384
constexpr int kU = 4;                      // independent 128-bit loads in flight per stream
constexpr int kNPart = 8;                  // doubles per task record (MHAQ_FQ_NPART)

struct Geom {
    int64_t n_rows, n_inner, n_ch;
    int64_t subs_per_row;    // ceil(n_inner / 4096)
    int64_t tasks_per_row;   // ceil(subs_per_row / spt)
    int64_t n_tasks;         // n_rows * tasks_per_row
    int64_t groups_per_row;  // ceil(tasks_per_row / kGroupTasks): first level of the ticketed reduction
    int64_t n_groups;        // n_rows * groups_per_row
    int spt;                 // sub-tiles per task
};

constexpr int kGroupTasks = 64;

// Task-size policy.  Measured on B200 (profiles/r01_task_granularity.txt): with one CTA per
// task and the hardware block scheduler balancing the SMs, SMALL tasks win — a lone CTA
// sustains only ~16 GB/s, so big tasks leave a long under-subscribed tail.
//   streaming kernels (forward, noise, stats): 1 sub-tile (4096 elements) per task
//   backward: 2 sub-tiles per task when the tensor is big enough, halving the per-task
//             flush (warp shuffles + ticket) cost per element
enum { GEOM_STREAM = 0, GEOM_REDUCE = 1 };

// Pure function of the shape (and policy): every kernel of a pass agrees on it, which is
// what makes the workspace layout part of the ABI.
__host__ __device__ inline Geom make_geom(int64_t n_rows, int64_t n_inner, int64_t n_ch, int policy,
                                          int spt_override = 0) {
    Geom g;
    g.n_rows = n_rows;
    g.n_inner = n_inner;
    g.n_ch = n_ch < 1 ? 1 : n_ch;
    g.subs_per_row = (n_inner + kSubElems - 1) / kSubElems;
    const int64_t total = n_rows * g.subs_per_row;
    int spt = 1;
    // 2 sub-tiles per task once that still leaves >= 4 waves of CTAs (148 SMs x 8 resident)
    if (policy == GEOM_REDUCE && g.subs_per_row >= 2 && total >= 2 * 4736) spt = 2;
    if (spt_override > 0) {
        spt = spt_override;
        while (spt > 1 && spt > g.subs_per_row) spt >>= 1;
    }
    g.spt = spt;
    g.tasks_per_row = (g.subs_per_row + spt - 1) / spt;
    g.n_tasks = n_rows * g.tasks_per_row;
    g.groups_per_row = (g.tasks_per_row + kGroupTasks - 1) / kGroupTasks;
    g.n_groups = n_rows * g.groups_per_row;
    return g;
}

// How the four parameter pointers are interpreted (MHAQ_FQ_PARAMS_* in the header).
//   LINEAR : scale, zero_point, min_val, max_val as given (strides 0/1)
//   ACT_LOG: the three NoisyAct parameters (gdnsq_act.py:42-48), each one float:
//              scale -> log_act_s,  zp -> act_b,  lo -> log_act_q,  hi unused
//            s = exp2(log_act_s), q = exp2(log_act_q), zp = lo = act_b, hi = (act_b + q) - s
//   WEIGHT_LOG: scale -> log_wght_s[ch] (stride 0/1), zp -> zero point (row minimum), no clamp
//            s = exp2(log_wght_s)                               (gdnsq_conv2d.py:72, 80-84)
// exp2f is the same libdevice routine torch's CUDA exp2 kernel calls, so the scale has the
// same bits as torch.exp2(log_s) (checked by tests/test_gpu_layers.py).
//   UNIT   : the value is ALREADY scaled (the reference's two-step form `v + QN*.apply(v, s)`,
//            gdnsq.py:204-208): s = 1, zp = 0, no clamp; `scale` only fixes the channel layout.
//            The finalize returns the estimator's own scale gradient (the noise / LSQ term).
enum { PARAMS_LINEAR = 0, PARAMS_ACT_LOG = 1, PARAMS_WEIGHT_LOG = 2, PARAMS_UNIT = 3 };

struct QParams {
    const float *scale, *zp, *lo, *hi;
    int ss, zs, ls, hs;
    int mode;
    // optional: gradients that reach the same four parameters along OTHER paths of the caller's
    // graph (one value per channel, NULL = none).  The reduction's last thread adds them to the
    // gradients it emits, so autograd sees ONE gradient per parameter and launches no
    // accumulation kernel (mhaq_fq_bwd_fused_acc_f32).
    const float *acc0, *acc1, *acc2, *acc3;
};

// Per-channel constants held in registers for the lifetime of a task.
struct QConst {
    float s, zp, lo, hi;
};

__device__ __forceinline__ QConst load_qconst(const QParams &p, int64_t ch) {
    QConst q;
    if (p.mode == PARAMS_ACT_LOG) {
        const float b = __ldg(p.zp);
        q.s = exp2f(__ldg(p.scale));
        q.zp = b;
        q.lo = b;
        q.hi = __fsub_rn(__fadd_rn(b, exp2f(__ldg(p.lo))), q.s);      // (act_b + q) - s
        return q;
    }
    if (p.mode == PARAMS_UNIT) {
        q.s = 1.f; q.zp = 0.f; q.lo = -INFINITY; q.hi = INFINITY;
        return q;
    }
    const float sv = __ldg(p.scale + ch * p.ss);
    q.s = (p.mode == PARAMS_WEIGHT_LOG) ? exp2f(sv) : sv;
    q.zp = __ldg(p.zp + ch * p.zs);
    q.lo = p.lo ? __ldg(p.lo + ch * p.ls) : -INFINITY;
    q.hi = p.hi ? __ldg(p.hi + ch * p.hs) : INFINITY;
    return q;
}

// ---- memory access ---------------------------------------------------------
// Streaming 128-bit load: read-only path, no L1 allocation (each element is
// touched exactly once per kernel).
__device__ __forceinline__ float4 ld_stream4(const float *p) {
    float4 v;
#ifndef MHAQ_LD_VOLATILE
#define MHAQ_LD_VOLATILE volatile
#endif
    asm MHAQ_LD_VOLATILE("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream4(float *p, const float4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// Load the 4 elements a thread owns at row position p (p % 4 == 0).
// VEC: one 128-bit access (requires 16-byte aligned row bases and n_inner % 4 == 0).
// !VEC: four predicated scalar accesses (ragged / unaligned rows).
template <bool VEC>
__device__ __forceinline__ float4 load4(const float *row, int64_t p, int64_t n_inner) {
    if (VEC) {
        return ld_stream4(row + p);
    } else {
        float4 v;
        v.x = (p + 0 < n_inner) ? __ldg(row + p + 0) : 0.f;
        v.y = (p + 1 < n_inner) ? __ldg(row + p + 1) : 0.f;
        v.z = (p + 2 < n_inner) ? __ldg(row + p + 2) : 0.f;
        v.w = (p + 3 < n_inner) ? __ldg(row + p + 3) : 0.f;
        return v;
    }
}
template <bool VEC>
__device__ __forceinline__ void store4(float *row, int64_t p, int64_t n_inner, const float4 &v) {
    if (VEC) {
        st_stream4(row + p, v);
    } else {
        if (p + 0 < n_inner) row[p + 0] = v.x;
        if (p + 1 < n_inner) row[p + 1] = v.y;
        if (p + 2 < n_inner) row[p + 2] = v.z;
        if (p + 3 < n_inner) row[p + 3] = v.w;
    }
}
// number of valid elements (0..4) of the float4 at row position p
template <bool VEC>
__device__ __forceinline__ int valid4(int64_t p, int64_t n_inner) {
    if (VEC) return p < n_inner ? 4 : 0;
    int64_t r = n_inner - p;
    return r <= 0 ? 0 : (r >= 4 ? 4 : (int)r);
}

// ---- arithmetic contract ---------------------------------------------------
// Every step below is a separate IEEE-754 fp32 round-to-nearest operation —
// exactly what the reference's chain of ATen kernels does
// (gdnsq.py:197-208, 229).  The file is compiled with --fmad=false and these
// wrappers use the _rn intrinsics so no FMA contraction can sneak in.
__device__ __forceinline__ float f_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float f_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float f_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float f_div(float a, float b) { return __fdiv_rn(a, b); }

// ---- exact division by a per-channel-invariant divisor ----------------------
// The hot loops divide by the same scale `s` for every element of a channel, so
// the reciprocal y = RN(1/s) is computed once (IEEE, __frcp_rn) and each
// quotient is one multiply plus one Markstein correction, the remainder
// r = a - q*s being evaluated EXACTLY by an FMA:
//     q0 = RN(a*y);   q1 = RN(q0 + r0*y)   ==   RN(a/s)
// for every a, s whose quotient and remainder stay in the normal range.
// tools/verify_fastdiv.cu checks this (and the two-correction variant nvcc itself
// emits for `/`) against __fdiv_rn over ALL 2^46 (significand(a), significand(s))
// pairs plus the exponent corners of the guards: zero mismatches
// (profiles/r01_fastdiv_exhaustive.txt).  3 instructions instead of ~10
// (MUFU.RCP + 6 FFMA + FCHK + BSSY/BRA/BSYNC) per division.
// The remainder is formed as q*s - a and negated on use, which keeps the sign of
// a zero quotient (-0/s = -0) without extra instructions.  NaN propagates.
// Out-of-range operands are the caller's job (see the guards in mhaq_fq.cu).
__device__ __forceinline__ float div_exact(float a, float s, float y) {
    const float q = __fmul_rn(a, y);
    const float r = __fmaf_rn(q, s, -a);
    return __fmaf_rn(-r, y, q);
}
// Two corrections (the textbook IEEE sequence); only used by the verifier.
__device__ __forceinline__ float div_exact2(float a, float s, float y) {
    float q = __fmul_rn(a, y);
    float r = __fmaf_rn(q, s, -a);
    q = __fmaf_rn(-r, y, q);
    r = __fmaf_rn(q, s, -a);
    return __fmaf_rn(-r, y, q);
}
// gu = RN(gv/s) when gv = RN(go*s): go itself is a faithful estimate of the
// quotient, so ONE exact-remainder correction from it is correctly rounded
// (checked exhaustively, same tool).  2 FMAs instead of a division.
__device__ __forceinline__ float div_of_product(float go, float gv, float s, float y) {
    const float r = __fmaf_rn(go, s, -gv);
    return __fmaf_rn(-r, y, go);
}
// Ranges inside which the FMA sequences above are exact (no intermediate
// under/overflow):  2^-32 <= s <= 2^32,  operand of a gradient division zero or
// 2^-56 <= |go| (no upper bound needed: see DESIGN.md), forward input |x| <= 2^80.
constexpr float kScaleLo = 2.3283064365386963e-10f;   // 2^-32
constexpr float kScaleHi = 4294967296.0f;             // 2^32
constexpr float kXHi = 1.2089258196146292e+24f;       // 2^80
constexpr uint32_t kGoLoBits2m1 = (0x23800000u << 1) - 1u;   // (bits(2^-56) << 1) - 1
constexpr float kGvHi = 1.2089258196146292e+24f;      // 2^80 (general gradient division)
__device__ __forceinline__ bool scale_fast_ok(float s) { return s >= kScaleLo && s <= kScaleHi; }

// ---- packed fp32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2) -----------------
// Blackwell issues IEEE round-to-nearest add/mul/fma on TWO fp32 lanes with one
// instruction (PTX add/sub/mul/fma.rn.f32x2 on a 64-bit register pair; scalar
// operands broadcast for free in SASS).  Each lane rounds exactly like the scalar
// instruction, so the arithmetic contract is unchanged while the hot loops — which are
// issue-slot limited, not pipe limited (profiles/r01_ncu_summary.md) — spend half the
// issue slots on floating-point work.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f32x2 bc2(float a) { return pk2(a, a); }
__device__ __forceinline__ void upk2(f32x2 v, float &lo, float &hi) {
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// A product whose result feeds an ADD must keep its own rounding (the reference rounds
// a*b before adding).  ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even
// though both carry an explicit rounding modifier (it does not do that to the scalar
// forms; it even rewrites fma2(a,b,-0) back into a fusable multiply) — found by the parity
// tests.  Such products are therefore formed with two scalar mul.rn.f32 and re-packed.
__device__ __forceinline__ f32x2 mul2_rounded(f32x2 a, f32x2 b) {
    float a0, a1, b0, b1;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(a0), "=f"(a1) : "l"(a));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(b0), "=f"(b1) : "l"(b));
    return pk2(__fmul_rn(a0, b0), __fmul_rn(a1, b1));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// Packed per-channel constants of the FMA division sequences.  Negated copies stand in
// for the operand-negate modifiers PTX does not expose on f32x2:
//   quotient  q' = fma(fma(q, -s, a), y, q)        with q = a*y   (== div_exact; the sign of a
//             ZERO quotient may differ, which is invisible after code = v + (rint(v) - v))
//   gradient  gu = fma(fma(go, s, -g), -y, go)     with -g = go*(-s) exactly (== div_of_product,
//             zero signs included: (+0)*(-y) = -0)
struct Div2 {
    f32x2 s, sn, y, yn;
};
__device__ __forceinline__ Div2 make_div2(float s, float y) {
    Div2 d;
    d.s = bc2(s); d.sn = bc2(-s); d.y = bc2(y); d.yn = bc2(-y);
    return d;
}
__device__ __forceinline__ f32x2 div2_quot(f32x2 a, const Div2 &d) {
    const f32x2 q = mul2(a, d.y);
    return fma2(fma2(q, d.sn, a), d.y, q);
}

// torch.clamp(x, lo, hi): NaN in x propagates; lo > hi yields hi.
__device__ __forceinline__ float f_clamp(float x, float lo, float hi) {
    float c = (x < lo) ? lo : x;
    return (c > hi) ? hi : c;
}

// ---- Philox4x32-10 ---------------------------------------------------------
struct PhiloxKey {
    uint32_t k0, k1;   // seed
    uint32_t o0, o1;   // offset (per-call stream id)
};

__host__ __device__ inline uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0;
        uint64_t p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    return make_uint4(c0, c1, c2, c3);
}

// 128 random bits for (row, super-tile T, thread tid).
__device__ __forceinline__ uint4 noise_block(const PhiloxKey &key, int64_t row,
                                             int64_t supers_per_row, int64_t T, int tid) {
    uint64_t pos = ((uint64_t)row * (uint64_t)supers_per_row + (uint64_t)T) * kThreads + tid;
    return philox4x32_10((uint32_t)pos, (uint32_t)(pos >> 32), key.o0, key.o1, key.k0, key.k1);
}
// the 4 bits of iteration `it` (0..31): bit k <-> element k of the float4
__device__ __forceinline__ uint32_t noise_nibble(const uint4 &b, int it) {
    uint32_t w = (it < 16) ? ((it < 8) ? b.x : b.y) : ((it < 24) ? b.z : b.w);
    return (w >> ((it & 7) * 4)) & 0xFu;
}

__device__ __forceinline__ PhiloxKey make_key(uint64_t seed, uint64_t offset,
                                              const uint64_t *philox_dev) {
    if (philox_dev) {
        seed = philox_dev[0];
        offset += philox_dev[1];
    }
    PhiloxKey k;
    k.k0 = (uint32_t)seed;
    k.k1 = (uint32_t)(seed >> 32);
    k.o0 = (uint32_t)offset;
    k.o1 = (uint32_t)(offset >> 32);
    return k;
}

// ---- reductions ------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Decompose a task id.  All threads of the CTA compute the same values.
struct Task {
    int64_t row, ch;
    int64_t q0, q1;      // sub-tile range within the row
    int64_t row_off;     // element offset of the row start
};
__device__ __forceinline__ Task make_task(const Geom &g, int64_t t) {
    Task k;
    int64_t j;
    if (g.tasks_per_row == 1) {
        k.row = t;
        j = 0;
    } else if (g.n_rows == 1) {
        k.row = 0;
        j = t;
    } else {
        k.row = t / g.tasks_per_row;
        j = t - k.row * g.tasks_per_row;
    }
    k.ch = (g.n_ch == 1) ? 0 : (k.row % g.n_ch);
    k.q0 = j * g.spt;
    k.q1 = k.q0 + g.spt;
    if (k.q1 > g.subs_per_row) k.q1 = g.subs_per_row;
    k.row_off = k.row * g.n_inner;
    return k;
}

}  // namespace mhaq
