// verify_fastdiv.cu — proof by exhaustion for the FMA division sequences of
// mhaq_b200/csrc/fq_common.cuh (div_exact, div_of_product; div_exact2 = the two-correction textbook sequence).
//
// Correct rounding of a quotient depends only on the two 24-bit significands
// (scaling either operand by a power of two is exact as long as nothing leaves the
// normal range, which the kernels' range guards ensure).  So checking
//     a = 1.ma in [1,2),  s = 1.ms in [1,2)   for ALL ma, ms in [0, 2^23)
// against __fdiv_rn covers every normal operand pair: 2^46 divisions per sequence.
// A sample of other exponents (incl. quotients near the guard limits) is run as well.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -I mhaq_b200/csrc \
//        -o tools/verify_fastdiv tools/verify_fastdiv.cu
//   ./tools/verify_fastdiv [ms_stride]      (1 = exhaustive, ~2 GPU-minutes on a B200)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "fq_common.cuh"

using namespace mhaq;

struct Counts {
    unsigned long long n, bad_exact, bad_near, bad_prod;
    unsigned int ex_a, ex_s;   // first failing pair (bits)
};

__global__ void check_kernel(uint32_t ms0, uint32_t n_ms, uint32_t ms_stride, int ea, int es, Counts *out) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_ms) return;
    const uint32_t ms = ms0 + idx * ms_stride;
    if (ms >= (1u << 23)) return;
    const float s = __uint_as_float(((uint32_t)(127 + es) << 23) | ms);
    const float y = __frcp_rn(s);
    unsigned long long be = 0, bn = 0, bp = 0;
    uint32_t fa = 0;
    for (uint32_t ma = 0; ma < (1u << 23); ++ma) {
        const float a = __uint_as_float(((uint32_t)(127 + ea) << 23) | ma);
        const float t = __fdiv_rn(a, s);
        const float q2 = div_exact(a, s, y);
        const float q1 = div_exact2(a, s, y);
        // product shortcut: go = a, gv = RN(go*s), want RN(gv/s)
        const float gv = __fmul_rn(a, s);
        const float tp = __fdiv_rn(gv, s);
        const float qp = div_of_product(a, gv, s, y);
        if (__float_as_uint(q2) != __float_as_uint(t)) { if (!be) fa = __float_as_uint(a); ++be; }
        if (__float_as_uint(q1) != __float_as_uint(t)) ++bn;
        if (__float_as_uint(qp) != __float_as_uint(tp)) { if (!bp && !be) fa = __float_as_uint(a); ++bp; }
    }
    atomicAdd(&out->n, (unsigned long long)(1u << 23));
    if (be) atomicAdd(&out->bad_exact, be);
    if (bn) atomicAdd(&out->bad_near, bn);
    if (bp) atomicAdd(&out->bad_prod, bp);
    if (be || bp) { out->ex_a = fa; out->ex_s = __float_as_uint(s); }
}

// special values: zeros keep their sign, NaN propagates
__global__ void special_kernel(int *fail) {
    const float s = 0.3f, y = __frcp_rn(s);
    float pz = 0.f, nz = -0.f, qnan = __uint_as_float(0x7fc00000u);
    if (__float_as_uint(div_exact(pz, s, y)) != 0x00000000u) atomicAdd(fail, 1);
    if (__float_as_uint(div_exact(nz, s, y)) != 0x80000000u) atomicAdd(fail, 1);
    if (__float_as_uint(div_of_product(pz, __fmul_rn(pz, s), s, y)) != 0x00000000u) atomicAdd(fail, 1);
    if (__float_as_uint(div_of_product(nz, __fmul_rn(nz, s), s, y)) != 0x80000000u) atomicAdd(fail, 1);
    float r = div_exact(qnan, s, y);
    if (r == r) atomicAdd(fail, 1);
}

static Counts run(uint32_t stride, int ea, int es, uint32_t limit_ms) {
    Counts *d, h = {};
    cudaMalloc(&d, sizeof(Counts));
    cudaMemset(d, 0, sizeof(Counts));
    const uint32_t total = ((1u << 23) + stride - 1) / stride;
    const uint32_t n_all = limit_ms ? (limit_ms < total ? limit_ms : total) : total;
    const uint32_t chunk = 1u << 16;     // threads per launch (keeps each launch ~1 s)
    for (uint32_t i0 = 0; i0 < n_all; i0 += chunk) {
        uint32_t n = n_all - i0 < chunk ? n_all - i0 : chunk;
        check_kernel<<<(n + 127) / 128, 128>>>(i0 * stride, n, stride, ea, es, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(2); }
    }
    cudaMemcpy(&h, d, sizeof(Counts), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return h;
}

int main(int argc, char **argv) {
    uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 1;
    if (stride < 1) stride = 1;
    int *dfail, hfail = 0;
    cudaMalloc(&dfail, sizeof(int));
    cudaMemset(dfail, 0, sizeof(int));
    special_kernel<<<1, 1>>>(dfail);
    cudaMemcpy(&hfail, dfail, sizeof(int), cudaMemcpyDeviceToHost);
    printf("special values (signed zeros, NaN): %s\n", hfail ? "FAIL" : "ok");
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    Counts c = run(stride, 0, 0, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("significand sweep: ms_stride=%u  pairs=%llu  (%.1f s)\n", stride, c.n, ms / 1e3);
    printf("  div_exact      (1 correction, used by the kernels) mismatches vs __fdiv_rn: %llu\n", c.bad_exact);
    printf("  div_of_product (gradient shortcut, used by the kernels) mismatches vs __fdiv_rn: %llu\n", c.bad_prod);
    printf("  div_exact2     (2 corrections, textbook)            mismatches vs __fdiv_rn: %llu\n", c.bad_near);
    if (c.bad_exact || c.bad_prod) printf("  first failing pair: a=0x%08x s=0x%08x\n", c.ex_a, c.ex_s);
    int bad = (c.bad_exact || c.bad_prod || hfail);
    // other exponents, sub-sampled significands: the guard-range corners
    // {exponent of a, exponent of s, check the product shortcut too (a = go must be >= 2^-56)}
    const int combos[][3] = {{-56, -32, 1}, {-56, 32, 1}, {-33, -32, 1}, {48, -32, 1}, {60, 32, 1},
                             {0, 31, 1},    {5, -7, 1},   {-20, 13, 1},  {-95, 32, 0}, {-95, -32, 0},
                             {90, -32, 1},  {94, 32, 1}};
    for (auto &cb : combos) {
        Counts k = run(4099, cb[0], cb[1], 0);
        printf("exponents ea=%d es=%d (ms_stride 4099): pairs=%llu exact_bad=%llu prod_bad=%llu%s\n", cb[0],
               cb[1], k.n, k.bad_exact, k.bad_prod, cb[2] ? "" : " (product shortcut outside its guard: not required)");
        bad |= (k.bad_exact || (cb[2] && k.bad_prod));
    }
    printf(bad ? "RESULT: FAIL\n" : "RESULT: PASS\n");
    return bad;
}
