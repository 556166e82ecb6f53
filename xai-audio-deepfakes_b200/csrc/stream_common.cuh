// Pieces shared by the streaming (barrier-free, warp-pipelined) overlap-add kernels: transform4_kernels.cu (n_fft 512) and
// transform5_kernels.cu (n_fft 1024) - mbarrier arrivals, MUFU forms, the per-frame mask gains, the unit list position.
#pragma once
#include "transform_common.cuh"
#include "fft3.cuh"

namespace adv {

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// the executing thread's earlier cp.async copies arrive on `bar` when they have landed (no pending-count increment:
// the barrier's expected count includes one such arrival per thread)
// non-blocking phase test (true once the phase of the given parity has completed)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// er = expm1(m log1p(a)) through MUFU (a >= 1/16) - see mask_gains() in transform_common.cuh for the error analysis
__device__ __forceinline__ float er_mufu(float a, float m) { return ex2_approx(m * lg2_approx(1.0f + a)) - 1.0f; }
__device__ __forceinline__ float er_series(float a, float m) {
    float L = fmaf(a, -1.0f / 6.0f, 0.2f);
    L = fmaf(-a, L, 0.25f);
    L = fmaf(-a, L, 1.0f / 3.0f);
    L = fmaf(-a, L, 0.5f);
    L = fmaf(-a, L, 1.0f);
    const float y = m * (L * a);
    float e = fmaf(y, 1.0f / 120.0f, 1.0f / 24.0f);
    e = fmaf(y, e, 1.0f / 6.0f);
    e = fmaf(y, e, 0.5f);
    e = fmaf(y, e, 1.0f);
    return y * e;
}

// One frame: one-sided spectrum x[9] (slot layout of fft3.cuh) and its mask column -> masked-in / masked-out spectra.
template <int MODE>
__device__ __forceinline__ void frame_gains(const float2* x, const float* m, float2* yr, float2* yi) {
    if constexpr (MODE == ADV_MASK_LINEAR) {
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const float gi = 1.0f - m[i];
            yr[i] = make_float2(x[i].x * m[i], x[i].y * m[i]);
            yi[i] = make_float2(x[i].x * gi, x[i].y * gi);
        }
    } else {
        float a[9], ia[9], amin = 1.0f;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const float r2 = fmaxf(fmaf(x[i].x, x[i].x, x[i].y * x[i].y), 1e-30f);
            ia[i] = rsqrt_ftz(r2);
            a[i] = r2 * ia[i];
            amin = fminf(amin, a[i]);
        }
        float er[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) er[i] = er_mufu(a[i], m[i]);
        if (__any_sync(0xffffffffu, amin < 0.0625f)) {  // (warp-uniform) some bin of the frame needs the series
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const float es = er_series(a[i], m[i]);
                er[i] = a[i] < 0.0625f ? es : er[i];
            }
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const float ei = (a[i] - er[i]) * rcp_ftz(1.0f + er[i]);
            const float gr = er[i] * ia[i], gi = ei * ia[i];
            yr[i] = make_float2(x[i].x * gr, x[i].y * gr);
            yi[i] = make_float2(x[i].x * gi, x[i].y * gi);
        }
    }
}

struct UnitPos {   // position of a unit in the batch-wide unit list
    int b, u;      // clip, unit inside the clip
    __device__ __forceinline__ void advance(int step, int upc) {
        u += step;
        while (u >= upc) { u -= upc; ++b; }
    }
};


}  // namespace adv
