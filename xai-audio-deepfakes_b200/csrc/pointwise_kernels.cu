// HBM-bound elementwise / reduction kernels of the path (sm_100a):
//   row_stats / normalize - zero_mean_unit_var_norm            classifier_embedder.py:59-63
//   lmac_reduce           - sigmoid + FF/Fid/AD/AI/AG + sums   LMAC_metrics.py:31-73,160-172
//   mask_apply            - standalone mask arithmetic         LMAC_metrics.py:136-143,151-153; loss_function.py:36-45
//   td_mask               - saliency time-domain mask          captum_saliency.py:136-143
//   mask_head             - 1x1 conv (C->1) + sigmoid          addvisor.py:57-60,82
//   band_swap             - complex row replacement            train_logReg_swapping.py:64-75, hifigan.py:206-214
// All are single-pass streaming kernels: coalesced 128-bit accesses where alignment allows, fp64
// partial sums reduced in a fixed order (bit-reproducible for a given launch shape).
#include <cstdlib>
#include "adv_internal.cuh"

namespace adv {

constexpr int kRowChunk = 8192;  // samples per row_stats block
constexpr int kPwThreads = 256;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- per-row partial sums ----------------------------------------------------------------------
__global__ void __launch_bounds__(kPwThreads)
row_stats_kernel(const float* __restrict__ in, int n, int parts, double* __restrict__ stats) {
    __shared__ double red[2][kPwThreads / 32];
    const int b = blockIdx.y, part = blockIdx.x;
    const float* row = in + (size_t)b * n;
    const int lo = part * kRowChunk, hi = min(n, lo + kRowChunk);
    double s = 0.0, ss = 0.0;
    for (int i = lo + threadIdx.x; i < hi; i += kPwThreads) {
        const float x = __ldg(row + i);
        s += (double)x;
        ss += (double)x * (double)x;
    }
    s = warp_sum_d(s);
    ss = warp_sum_d(ss);
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = s;
        red[1][threadIdx.x >> 5] = ss;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int k = 0; k < kPwThreads / 32; ++k) {
            a += red[0][k];
            c += red[1][k];
        }
        stats[((size_t)b * parts + part) * 2 + 0] = a;
        stats[((size_t)b * parts + part) * 2 + 1] = c;
    }
}

// 128-bit form (n % 4 == 0, 16-byte aligned rows): the 8 loads of a thread are in flight together; the same float64
// accumulation per thread, so the partial sums differ from the scalar form only in summation order
__global__ void __launch_bounds__(kPwThreads)
row_stats4_kernel(const float* __restrict__ in, int n, int parts, double* __restrict__ stats) {
    __shared__ double red[2][kPwThreads / 32];
    const int b = blockIdx.y, part = blockIdx.x;
    const float4* row = reinterpret_cast<const float4*>(in + (size_t)b * n);
    const int lo = part * (kRowChunk / 4), hi = min(n / 4, lo + kRowChunk / 4);
    constexpr int kPer = kRowChunk / 4 / kPwThreads;
    float4 v[kPer];
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
        const int i = lo + threadIdx.x + j * kPwThreads;
        v[j] = i < hi ? __ldg(row + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    double s = 0.0, ss = 0.0;
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
        const double x0 = v[j].x, x1 = v[j].y, x2 = v[j].z, x3 = v[j].w;
        s += (x0 + x1) + (x2 + x3);
        ss += (x0 * x0 + x1 * x1) + (x2 * x2 + x3 * x3);
    }
    s = warp_sum_d(s);
    ss = warp_sum_d(ss);
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = s;
        red[1][threadIdx.x >> 5] = ss;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int k = 0; k < kPwThreads / 32; ++k) {
            a += red[0][k];
            c += red[1][k];
        }
        stats[((size_t)b * parts + part) * 2 + 0] = a;
        stats[((size_t)b * parts + part) * 2 + 1] = c;
    }
}

// ---- (x - mean) / (std_unbiased + 1e-7) -----------------------------------------------------------
// blockIdx.z selects one of up to two arrays that share a stats buffer (the explain kernel's rel / irr
// outputs: (sum, sumsq) pairs at columns col and col + 2), so both normalisers are one launch.
// (<= 32 registers per thread: CTAs of this kernel fit next to a resident explain CTA of generation 2 / 3 - 96
//  registers x 512 threads - so the normaliser of batch i streams through L2 while the issue-bound explain kernel of
//  batch i+1 owns the issue slots.  The generation-4 explain kernel takes all registers of an SM, see its register cap.)
__device__ __forceinline__ float sigmoidf_ref(float x) { return 1.0f / (1.0f + expf(-x)); }

// The metric reduction of the SAME batch as one extra CTA of the normaliser's grid (up to 1 024 clips): the two are
// independent, and as separate launches the 4 us reduction held an SM - and with it the start of the next batch's
// persistent explain CTA on that SM - after the normaliser had finished (LMAC_metrics.py:31-73,160-172; same arithmetic
// and summation order as lmac_kernel's single-block path).
struct LmacJob {
    const float* p;
    const float* th;
    const float* q;
    int n, flags;
    float* scores;
    double* sums;
};
__device__ __forceinline__ void lmac_single_block(const LmacJob& j, double (*red)[kPwThreads / 32]) {
    const bool is_logit = j.flags & ADV_LMAC_LOGITS, accumulate = j.flags & ADV_LMAC_ACCUMULATE;
    double acc[5] = {0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < j.n; i += kPwThreads) {
        float p = j.p[i], th = j.th[i], q = j.q[i];
        if (is_logit) {
            p = sigmoidf_ref(p);
            th = sigmoidf_ref(th);
            q = sigmoidf_ref(q);
        }
        const float pc = (p > 0.5f) ? p : 1.0f - p;
        const float oc = (th > 0.5f) ? th : 1.0f - th;
        const float d = p - 0.5f;
        const float sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
        const float ff = (p - q) * sgn;
        const float fid = ((p > 0.5f) == (th > 0.5f)) ? 1.f : 0.f;
        const float ad = (fmaxf(pc - oc, 0.f) / (pc + 1e-10f)) * 100.f;
        const float ai = (oc > pc) ? 100.f : 0.f;
        const float ag = (fmaxf(oc - pc, 0.f) / ((1.f - pc) + 1e-10f)) * 100.f;
        if (j.scores != nullptr) {
            float* s = j.scores + (size_t)i * 7;
            s[0] = ff; s[1] = fid; s[2] = ad; s[3] = ai; s[4] = ag; s[5] = pc; s[6] = oc;
        }
        acc[0] += ff; acc[1] += fid; acc[2] += ad; acc[3] += ai; acc[4] += ag;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        acc[k] = warp_sum_d(acc[k]);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = acc[k];
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double s = 0.0;
        for (int w = 0; w < kPwThreads / 32; ++w) s += red[threadIdx.x][w];
        s = 0.0 + s;
        j.sums[threadIdx.x] = (accumulate ? j.sums[threadIdx.x] : 0.0) + s;
    }
    if (threadIdx.x == 5) j.sums[5] = (accumulate ? j.sums[5] : 0.0) + (double)j.n;
}

#ifndef ADV_NORM_PRELOAD
#define ADV_NORM_PRELOAD 1
#endif
template <int NT>
__global__ void __launch_bounds__(NT, 2048 / NT)
normalize_kernel(const float* __restrict__ in0, float* __restrict__ out0, const float* __restrict__ in1,
                 float* __restrict__ out1, int n, const double* __restrict__ stats, int parts, int width, int col,
                 LmacJob job) {
    __shared__ float s_mean, s_den;
    pdl_launch_dependents();
    pdl_wait();
    if (NT == kPwThreads && job.p != nullptr && blockIdx.x == gridDim.x - 1) {   // the grid's extra column: the metric CTA
        if (blockIdx.y == 0 && blockIdx.z == 0) {
            __shared__ double red[5][kPwThreads / 32];
            lmac_single_block(job, red);
        }
        return;
    }
    const int b = blockIdx.y, z = blockIdx.z;
    const float* in = z ? in1 : in0;
    float* out = z ? out1 : out0;
    const float* row = in + (size_t)b * n;
    float* orow = out + (size_t)b * n;
    const int lo = blockIdx.x * kRowChunk, hi = min(n, lo + kRowChunk);
    const bool vec = ((n & 3) == 0) && (((uintptr_t)in | (uintptr_t)out) & 15) == 0;
    const float4* r4 = reinterpret_cast<const float4*>(row);
    float4* o4 = reinterpret_cast<float4*>(orow);
    constexpr int U = 4;
    // the first samples of the chunk are requested before the statistics are folded: the whole grid is one wave, so the
    // fold's two dependent round trips (partials, then a double-precision divide and square root) would otherwise pass
    // with no load in flight anywhere
    float4 v[U];
    int i = lo / 4 + threadIdx.x;
    if (threadIdx.x >= 32) {   // (warp 0 folds first: it has no registers to hold samples across the fold)
#if ADV_NORM_PRELOAD
        if (vec) {
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i + u * NT < hi / 4) v[u] = __ldg(r4 + i + u * NT);
        }
#endif
    } else {  // warp 0 folds the per-tile partials in a fixed order (lane-strided + butterfly)
        double s = 0.0, ss = 0.0;
        const double* st = stats + (size_t)b * parts * width + col + 2 * z;
        for (int k = threadIdx.x; k < parts; k += 32) {
            s += st[(size_t)k * width];
            ss += st[(size_t)k * width + 1];
        }
        s = warp_sum_d(s);
        ss = warp_sum_d(ss);
        if (threadIdx.x == 0) {
            const double mean = s / (double)n;
            double var = (ss - s * mean) / (double)(n - 1);  // unbiased (torch.std default)
            if (var < 0.0) var = 0.0;
            s_mean = (float)mean;
            s_den = (float)sqrt(var) + 1e-7f;
        }
    }
    __syncthreads();
    // one reciprocal per row instead of a division per sample (<= 1.5 ulp from the reference's quotient; the
    // IEEE division was a third of this kernel's instructions)
    const float mean = s_mean, inv = 1.0f / s_den;
#if ADV_NORM_PRELOAD
    if (vec && threadIdx.x < 32) {
#else
    if (vec) {
#endif
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i + u * NT < hi / 4) v[u] = __ldg(r4 + i + u * NT);
    }
    if (vec) {
        for (; i < hi / 4; i += NT * U) {
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i + u * NT < hi / 4) {
                    float4 y;
                    y.x = (v[u].x - mean) * inv;
                    y.y = (v[u].y - mean) * inv;
                    y.z = (v[u].z - mean) * inv;
                    y.w = (v[u].w - mean) * inv;
                    o4[i + u * NT] = y;
                }
            const int nx = i + NT * U;   // next round's loads
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (nx + u * NT < hi / 4) v[u] = __ldg(r4 + nx + u * NT);
        }
    } else {
        for (int j = lo + threadIdx.x; j < hi; j += NT) orow[j] = (__ldg(row + j) - mean) * inv;
    }
}

// ---- LMAC scores + deterministic two-level sum -----------------------------------------------------
constexpr int kLmacPerBlock = 1024;


__global__ void __launch_bounds__(kPwThreads, 8)
lmac_kernel(const float* __restrict__ p_in, const float* __restrict__ th_in, const float* __restrict__ q_in, int n,
            int flags, float* __restrict__ scores, double* __restrict__ sums, double* __restrict__ partials,
            unsigned int* __restrict__ counter) {
    const bool is_logit = flags & ADV_LMAC_LOGITS, accumulate = flags & ADV_LMAC_ACCUMULATE;
    pdl_launch_dependents();
    pdl_wait();
    __shared__ double red[5][kPwThreads / 32];
    __shared__ bool last;
    double acc[5] = {0, 0, 0, 0, 0};
    const int lo = blockIdx.x * kLmacPerBlock, hi = min(n, lo + kLmacPerBlock);
    for (int i = lo + threadIdx.x; i < hi; i += kPwThreads) {
        float p = p_in[i], th = th_in[i], q = q_in[i];
        if (is_logit) {  // classifier_embedder.py:36
            p = sigmoidf_ref(p);
            th = sigmoidf_ref(th);
            q = sigmoidf_ref(q);
        }
        // LMAC_metrics.py:43-45 (pred*p + (1-pred)*(1-p) is exactly p or 1-p)
        const float pc = (p > 0.5f) ? p : 1.0f - p;
        const float oc = (th > 0.5f) ? th : 1.0f - th;
        const float d = p - 0.5f;
        const float sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
        const float ff = (p - q) * sgn;                                        // :48-52
        const float fid = ((p > 0.5f) == (th > 0.5f)) ? 1.f : 0.f;             // :31-38
        const float ad = (fmaxf(pc - oc, 0.f) / (pc + 1e-10f)) * 100.f;        // :55-59
        const float ai = (oc > pc) ? 100.f : 0.f;                              // :62-66
        const float ag = (fmaxf(oc - pc, 0.f) / ((1.f - pc) + 1e-10f)) * 100.f;  // :69-73
        if (scores != nullptr) {
            float* s = scores + (size_t)i * 7;
            s[0] = ff; s[1] = fid; s[2] = ad; s[3] = ai; s[4] = ag; s[5] = pc; s[6] = oc;
        }
        acc[0] += ff; acc[1] += fid; acc[2] += ad; acc[3] += ai; acc[4] += ag;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        acc[k] = warp_sum_d(acc[k]);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = acc[k];
    }
    __syncthreads();
    if (gridDim.x == 1) {
        // up to 1 024 clips (every per-batch call of the pipeline): the block's own fold IS the result - no partials, no
        // fence / counter / re-read round trips (the kernel was 5.6 us of dependent global accesses for 64 clips, and a
        // persistent explain CTA waiting for the SM it sits on starts that much later).  Same additions, same order.
        if (threadIdx.x < 5) {
            double s = 0.0;
            for (int w = 0; w < kPwThreads / 32; ++w) s += red[threadIdx.x][w];
            s = 0.0 + s;
            sums[threadIdx.x] = (accumulate ? sums[threadIdx.x] : 0.0) + s;
        }
        if (threadIdx.x == 5) sums[5] = (accumulate ? sums[5] : 0.0) + (double)n;
        return;
    }
    if (threadIdx.x < 5) {
        double s = 0.0;
        for (int w = 0; w < kPwThreads / 32; ++w) s += red[threadIdx.x][w];
        partials[(size_t)blockIdx.x * 5 + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (last) {  // the last block to finish folds the block partials in block order
        __threadfence();
        if (threadIdx.x < 5) {
            double s = 0.0;
            for (unsigned int k = 0; k < gridDim.x; ++k) s += partials[(size_t)k * 5 + threadIdx.x];
            sums[threadIdx.x] = (accumulate ? sums[threadIdx.x] : 0.0) + s;
        }
        if (threadIdx.x == 5) sums[5] = (accumulate ? sums[5] : 0.0) + (double)n;
        if (threadIdx.x == 0) *counter = 0u;  // re-arm for the next launch
    }
}

// ---- standalone mask-apply on (mag, phase) ---------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kPwThreads)
mask_apply_kernel(const float* __restrict__ mag, const float* __restrict__ phase, const float* __restrict__ mask,
                  int T, int F, int Fm, int Tm, float2* __restrict__ rel, float2* __restrict__ irr) {
    // 32(f) x 32(t) tile; the mask is [Fm][Tm] (t fastest), the spectra are frame-major (f fastest):
    // transpose the mask tile through shared memory so both sides are coalesced
    __shared__ float tile[32][33];
    const int b = blockIdx.z, f0 = blockIdx.x * 32, t0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const float* mrow = mask + (size_t)b * Fm * Tm;
    // every load of the thread (4 mask-tile elements, 4 magnitudes, 4 phases) is issued before the first use: one
    // memory latency per CTA instead of one per element
    float mreg[4], areg[4], preg[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int r = ty + 8 * j;
        const int fm = f0 + r, tm = t0 + tx;
        mreg[j] = (fm < Fm && tm < Tm) ? __ldg(mrow + (size_t)fm * Tm + tm) : 0.0f;
        const int t = t0 + r, f = f0 + tx;
        const bool ok = t < T && f < F;
        const size_t idx = ((size_t)b * T + (ok ? t : 0)) * F + (ok ? f : 0);
        areg[j] = ok ? __ldg(mag + idx) : 0.0f;
        preg[j] = ok ? __ldg(phase + idx) : 0.0f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) tile[ty + 8 * j][tx] = mreg[j];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int r = ty + 8 * j;
        const int t = t0 + r, f = f0 + tx;
        if (t >= T || f >= F) continue;
        const size_t idx = ((size_t)b * T + t) * F + f;
        const float m = tile[tx][r];
        const float a = areg[j];
        float s, c;
        // exp(1j*phase): angles come from atan2, |phase| <= pi, where the MUFU forms are accurate to 2^-21 absolute;
        // anything else (callers may pass unwrapped phases) takes libm's range reduction
        const float ph = preg[j];
        if (fabsf(ph) <= 3.2f) __sincosf(ph, &s, &c);
        else sincosf(ph, &s, &c);
        float ar, ai;
        if (MODE == ADV_MASK_LINEAR) {
            ar = m * a;
            ai = (1.0f - m) * a;
        } else {
            // the fused kernel's gains (transform_kernels.cu mask_gains): (1+a)^m - 1 through MUFU lg2 / ex2 for a >= 1/16,
            // Taylor series below; the complementary term needs no second exponential:
            // expm1((1-m) log1p a) = (1+a)/(1+er) - 1 = (a - er)/(1 + er).  libm's log1pf + 2 x expm1f made this kernel
            // instruction-bound (58.6 us for 64 clips = 48 % of the HBM roofline)
            if (a >= 0.0625f) {
                float lg, ex;
                asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(1.0f + a));
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(m * lg));
                ar = ex - 1.0f;
            } else {
                float L = fmaf(a, -1.0f / 6.0f, 0.2f);
                L = fmaf(-a, L, 0.25f);
                L = fmaf(-a, L, 1.0f / 3.0f);
                L = fmaf(-a, L, 0.5f);
                L = fmaf(-a, L, 1.0f);
                const float y = m * L * a;
                float e = fmaf(y, 1.0f / 120.0f, 1.0f / 24.0f);
                e = fmaf(y, e, 1.0f / 6.0f);
                e = fmaf(y, e, 0.5f);
                e = fmaf(y, e, 1.0f);
                ar = y * e;
            }
            ai = __fdividef(a - ar, 1.0f + ar);
        }
        rel[idx] = make_float2(ar * c, ar * s);
        irr[idx] = make_float2(ai * c, ai * s);
    }
}

// ---- time-domain saliency mask --------------------------------------------------------------------
__global__ void __launch_bounds__(kPwThreads)
rowmax_abs_kernel(const float* __restrict__ attr, int n, float* __restrict__ rowmax) {
    const int b = blockIdx.y;
    const float* row = attr + (size_t)b * n;
    const int lo = blockIdx.x * kRowChunk, hi = min(n, lo + kRowChunk);
    float m = 0.f;
    for (int i = lo + threadIdx.x; i < hi; i += kPwThreads) m = fmaxf(m, fabsf(__ldg(row + i)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    // non-negative floats order like their bit patterns: integer atomicMax is exact and order-free
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(rowmax + b), __float_as_int(m));
}
__global__ void __launch_bounds__(kPwThreads)
td_mask_kernel(const float* __restrict__ wave, const float* __restrict__ attr, int n,
               const float* __restrict__ rowmax, float* __restrict__ mask_out, float* __restrict__ rel,
               float* __restrict__ irr) {
    const int b = blockIdx.y;
    const float den = rowmax[b] + 1e-8f;
    const size_t base = (size_t)b * n;
    const int lo = blockIdx.x * kRowChunk, hi = min(n, lo + kRowChunk);
    for (int i = lo + threadIdx.x; i < hi; i += kPwThreads) {
        const float m = fabsf(__ldg(attr + base + i)) / den;
        const float w = __ldg(wave + base + i);
        if (mask_out != nullptr) mask_out[base + i] = m;
        rel[base + i] = w * m;
        irr[base + i] = w * (1.0f - m);
    }
}

// 128-bit forms (n % 4 == 0, 16-byte aligned rows): a thread's 8 loads per array are all in flight at once
__global__ void __launch_bounds__(kPwThreads)
rowmax_abs4_kernel(const float* __restrict__ attr, int n, float* __restrict__ rowmax) {
    const int b = blockIdx.y;
    const float4* row = reinterpret_cast<const float4*>(attr + (size_t)b * n);
    const int lo = blockIdx.x * (kRowChunk / 4), hi = min(n / 4, lo + kRowChunk / 4);
    constexpr int kPer = kRowChunk / 4 / kPwThreads;
    float4 v[kPer];
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
        const int i = lo + threadIdx.x + j * kPwThreads;
        v[j] = i < hi ? __ldg(row + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float m = 0.f;
#pragma unroll
    for (int j = 0; j < kPer; ++j)
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v[j].x), fabsf(v[j].y))), fmaxf(fabsf(v[j].z), fabsf(v[j].w)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(rowmax + b), __float_as_int(m));
}
__global__ void __launch_bounds__(kPwThreads)
td_mask4_kernel(const float* __restrict__ wave, const float* __restrict__ attr, int n,
                const float* __restrict__ rowmax, float* __restrict__ mask_out, float* __restrict__ rel,
                float* __restrict__ irr) {
    const int b = blockIdx.y;
    const float den = rowmax[b] + 1e-8f;
    const size_t base = (size_t)b * n / 4;
    const float4* a4 = reinterpret_cast<const float4*>(attr) + base;
    const float4* w4 = reinterpret_cast<const float4*>(wave) + base;
    float4* r4 = reinterpret_cast<float4*>(rel) + base;
    float4* i4 = reinterpret_cast<float4*>(irr) + base;
    float4* m4 = mask_out ? reinterpret_cast<float4*>(mask_out) + base : nullptr;
    const int lo = blockIdx.x * (kRowChunk / 4), hi = min(n / 4, lo + kRowChunk / 4);
    constexpr int kPer = kRowChunk / 4 / kPwThreads / 2;  // two rounds of 4 + 4 loads: 32 data registers in flight
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float4 av[kPer], wv[kPer];
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
            const int i = lo + threadIdx.x + (h * kPer + j) * kPwThreads;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            av[j] = i < hi ? __ldg(a4 + i) : z;
            wv[j] = i < hi ? __ldg(w4 + i) : z;
        }
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
            const int i = lo + threadIdx.x + (h * kPer + j) * kPwThreads;
            if (i >= hi) continue;
            // (the reference divides: abs(attr) / (max + 1e-8), captum_saliency.py:139-140 - kept as an IEEE division)
            const float4 m = make_float4(fabsf(av[j].x) / den, fabsf(av[j].y) / den, fabsf(av[j].z) / den,
                                         fabsf(av[j].w) / den);
            const float4 w = wv[j];
            if (m4 != nullptr) m4[i] = m;
            r4[i] = make_float4(w.x * m.x, w.y * m.y, w.z * m.z, w.w * m.w);
            i4[i] = make_float4(w.x * (1.0f - m.x), w.y * (1.0f - m.y), w.z * (1.0f - m.z), w.w * (1.0f - m.w));
        }
    }
}

// ---- mask head: sigmoid(sum_c w_c * y1[b,c,:] + bias) ------------------------------------------------
template <int C, bool VEC>
__global__ void __launch_bounds__(kPwThreads)
mask_head_kernel(const float* __restrict__ y1, const float* __restrict__ w, const float* __restrict__ bias,
                 int64_t hw, float* __restrict__ mask) {
    __shared__ float ws[C];
    if (threadIdx.x < C) ws[threadIdx.x] = w[threadIdx.x];
    __syncthreads();
    const int b = blockIdx.y;
    const float bb = bias[0];
    const float* src = y1 + (size_t)b * C * hw;
    float* dst = mask + (size_t)b * hw;
    if (VEC) {
        const int64_t nv = hw / 4;
        for (int64_t i = (int64_t)blockIdx.x * kPwThreads + threadIdx.x; i < nv; i += (int64_t)gridDim.x * kPwThreads) {
            float4 v[C];
#pragma unroll
            for (int c = 0; c < C; ++c) v[c] = __ldg(reinterpret_cast<const float4*>(src + (size_t)c * hw) + i);
            float4 a = make_float4(bb, bb, bb, bb);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                a.x = fmaf(ws[c], v[c].x, a.x);
                a.y = fmaf(ws[c], v[c].y, a.y);
                a.z = fmaf(ws[c], v[c].z, a.z);
                a.w = fmaf(ws[c], v[c].w, a.w);
            }
            a.x = sigmoidf_ref(a.x); a.y = sigmoidf_ref(a.y); a.z = sigmoidf_ref(a.z); a.w = sigmoidf_ref(a.w);
            reinterpret_cast<float4*>(dst)[i] = a;
        }
    } else {
        for (int64_t i = (int64_t)blockIdx.x * kPwThreads + threadIdx.x; i < hw; i += (int64_t)gridDim.x * kPwThreads) {
            float a = bb;
#pragma unroll
            for (int c = 0; c < C; ++c) a = fmaf(ws[c], __ldg(src + (size_t)c * hw + i), a);
            dst[i] = sigmoidf_ref(a);
        }
    }
}
__global__ void __launch_bounds__(kPwThreads)
mask_head_generic_kernel(const float* __restrict__ y1, const float* __restrict__ w, const float* __restrict__ bias,
                         int C, int64_t hw, float* __restrict__ mask) {
    const int b = blockIdx.y;
    const float bb = bias[0];
    const float* src = y1 + (size_t)b * C * hw;
    for (int64_t i = (int64_t)blockIdx.x * kPwThreads + threadIdx.x; i < hw; i += (int64_t)gridDim.x * kPwThreads) {
        float a = bb;
        for (int c = 0; c < C; ++c) a = fmaf(__ldg(w + c), __ldg(src + (size_t)c * hw + i), a);
        mask[(size_t)b * hw + i] = sigmoidf_ref(a);
    }
}

// ---- band swap -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPwThreads)
band_swap_kernel(const float2* __restrict__ real, const float2* __restrict__ voc, int64_t total, int F, int f_lo,
                 int f_hi, float2* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * kPwThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kPwThreads) {
        const int f = (int)(i % F);
        out[i] = (f >= f_lo && f < f_hi) ? __ldg(voc + i) : __ldg(real + i);
    }
}

// all bands of the fabrication loop at once (hifigan.py:206-214, train_logReg_swapping.py:64-75 run it once per
// 1 kHz band): out[k] = real with rows [edges[k], edges[k+1]) taken from voc; real / voc are read once per band
// from L2, the n_bands x larger output is what the one batched iSTFT that follows consumes
__global__ void __launch_bounds__(kPwThreads)
band_swap_multi_kernel(const float2* __restrict__ real, const float2* __restrict__ voc, int64_t total, int F,
                       const int* __restrict__ edges, int n_bands, float2* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * kPwThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kPwThreads) {
        const int f = (int)(i % F);
        const float2 r = __ldg(real + i), v = __ldg(voc + i);
        for (int k = 0; k < n_bands; ++k) out[(int64_t)k * total + i] = (f >= edges[k] && f < edges[k + 1]) ? v : r;
    }
}

// ---- gradient of the linear mask path w.r.t. the mask (backward of loss_function.py:36-47) -----------------
// 32 x 32 tiles: frame-major reads (f fastest) of X and A, transposed through shared memory so that the
// [B][Fm][Tm] gradient (t fastest) is written coalesced.
__global__ void __launch_bounds__(256)
mask_grad_linear_kernel(const float2* __restrict__ X, int64_t sb, int64_t st, int64_t sf, const float2* __restrict__ A,
                        int F, int T, int Fm, int Tm, float* __restrict__ gm) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, f0 = blockIdx.x * 32, t0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int t = t0 + ty + 8 * j, f = f0 + tx;
        float g = 0.0f;
        if (t < Tm && f < Fm) {
            const float2 x = __ldg(X + (size_t)b * sb + (size_t)t * st + (size_t)f * sf);
            const float2 a = __ldg(A + ((size_t)b * T + t) * F + f);
            const float c = (f == 0 || f == F - 1) ? 1.0f : 2.0f;
            g = c * fmaf(x.x, a.x, x.y * a.y);
        }
        tile[ty + 8 * j][tx] = g;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int f = f0 + ty + 8 * j, t = t0 + tx;
        if (f < Fm && t < Tm) gm[((size_t)b * Fm + f) * Tm + t] = tile[tx][ty + 8 * j];
    }
}

// ---- cross-correlation arg-max (align_waveforms, hifigan.py:113-136) ------------------------------------
// cc[j] = sum_i ref[j + i - P] * deg[i],  P = n_deg, j = 0 .. n_ref + P  (the reference pads ref with P zeros on both
// sides and runs conv1d); the alignment shift is argmax_j cc[j] - P.  Direct form, fp32 accumulation like the
// reference's conv1d: a CTA owns kXcLags consecutive lags, streams deg in chunks through shared memory and slides
// an 8-lag register window over the matching ref samples (1 shared load + 8 FMA per tap); chunks that only meet
// the zero padding are skipped.  Stage 2 folds the per-CTA maxima (lowest index wins ties).
constexpr int kXcThreads = 128, kXcPerThread = 8, kXcLags = kXcThreads * kXcPerThread, kXcChunk = 512;
__device__ __forceinline__ int xc_pad(int i) { return i + (i >> 5); }  // stride-8 accesses hit 32 different banks

__global__ void __launch_bounds__(kXcThreads)
xcorr_partial_kernel(const float* __restrict__ ref, int n_ref, const float* __restrict__ deg, int n_deg,
                     float* __restrict__ best_val, int* __restrict__ best_idx) {
    __shared__ float refs[kXcLags + kXcChunk + 8 + (kXcLags + kXcChunk + 8) / 32 + 1];
    __shared__ __align__(16) float degs[kXcChunk];
    __shared__ float rv[kXcThreads / 32];
    __shared__ int ri[kXcThreads / 32];
    const int P = n_deg, n_lags = n_ref + P + 1;
    const int j0 = blockIdx.x * kXcLags, t = threadIdx.x;
    float acc[kXcPerThread];
#pragma unroll
    for (int r = 0; r < kXcPerThread; ++r) acc[r] = 0.0f;
    // taps i that can meet a non-zero ref sample for some lag of this CTA: 0 <= j + i - P < n_ref
    const int i_lo = max(0, P - (j0 + kXcLags - 1)), i_hi = min(n_deg, n_ref + P - j0);
    for (int i0 = (i_lo / kXcChunk) * kXcChunk; i0 < i_hi; i0 += kXcChunk) {
        __syncthreads();
        for (int k = t; k < kXcChunk; k += kXcThreads) degs[k] = (i0 + k < n_deg) ? __ldg(deg + i0 + k) : 0.0f;
        const int base = j0 + i0 - P;  // ref index of window element 0
        for (int k = t; k < kXcLags + kXcChunk; k += kXcThreads) {
            const int idx = base + k;
            refs[xc_pad(k)] = (idx >= 0 && idx < n_ref) ? __ldg(ref + idx) : 0.0f;
        }
        __syncthreads();
        // 8 taps per step: the 8-lag window slides over 15 consecutive ref samples held in registers (8 carried over,
        // 8 new), the 8 taps arrive as two 128-bit broadcast loads: 64 FMA per 10 shared-memory instructions
        float w[2 * kXcPerThread];
#pragma unroll
        for (int r = 0; r < kXcPerThread; ++r) w[r] = refs[xc_pad(t * kXcPerThread + r)];
#pragma unroll 2
        for (int i = 0; i < kXcChunk; i += 8) {
            const float4 d0 = *reinterpret_cast<const float4*>(degs + i);
            const float4 d1 = *reinterpret_cast<const float4*>(degs + i + 4);
            const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
            for (int r = 0; r < kXcPerThread; ++r) w[kXcPerThread + r] = refs[xc_pad(t * kXcPerThread + kXcPerThread + i + r)];
#pragma unroll
            for (int tt = 0; tt < 8; ++tt)
#pragma unroll
                for (int r = 0; r < kXcPerThread; ++r) acc[r] = fmaf(w[r + tt], d[tt], acc[r]);
#pragma unroll
            for (int r = 0; r < kXcPerThread; ++r) w[r] = w[kXcPerThread + r];
        }
    }
    float bv = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int r = 0; r < kXcPerThread; ++r) {
        const int j = j0 + t * kXcPerThread + r;
        if (j < n_lags && (acc[r] > bv)) { bv = acc[r]; bi = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((t & 31) == 0) { rv[t >> 5] = bv; ri[t >> 5] = bi; }
    __syncthreads();
    if (t == 0) {
        for (int k = 1; k < kXcThreads / 32; ++k)
            if (rv[k] > bv || (rv[k] == bv && ri[k] < bi)) { bv = rv[k]; bi = ri[k]; }
        best_val[blockIdx.x] = bv;
        best_idx[blockIdx.x] = bi;
    }
}

__global__ void __launch_bounds__(256)
xcorr_final_kernel(const float* __restrict__ best_val, const int* __restrict__ best_idx, int n, int pad, int* __restrict__ shift) {
    __shared__ float rv[8];
    __shared__ int ri[8];
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int k = threadIdx.x; k < n; k += 256) {
        const float v = best_val[k];
        const int i = best_idx[k];
        if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { rv[threadIdx.x >> 5] = bv; ri[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k)
            if (rv[k] > bv || (rv[k] == bv && ri[k] < bi)) { bv = rv[k]; bi = ri[k]; }
        *shift = bi - pad;
    }
}

// ---- cross-correlation through the frequency domain (overlap-save, 1024-point frames, hop 512) ------------------
// With D[t] = zero-padded STFT frames of the 512-tap blocks of deg (rectangular 512-window centred in the frame) and
// R[t] = STFT frames of the zero-padded ref (full rectangular window), both frame-major [frames][513]:
//     Z[q + 1][k] = (-1)^k * sum_b conj(D[b + 1][k]) * R[q + b + 1][k],      Z[0][k] = 0
// is the spectrum whose iSTFT (512-window, hop 512) holds cc[512 q + j] at output sample 512 q + j + 256: the two
// quarter-frame shifts (block at in-frame offset 256 on the way in, valid lags moved under the synthesis window on
// the way out) are e^{-i pi k / 2} each.  One CTA per output frame, one thread per bin, 4 blocks of deg in flight.
__global__ void __launch_bounds__(256)
xcorr_fd_mac_kernel(const float2* __restrict__ D, int nb, const float2* __restrict__ R, int t_r, float2* __restrict__ Z,
                    int bins) {
    const int q = blockIdx.x;  // output frame q (frame 0 is the zero frame)
    for (int k = threadIdx.x; k < bins; k += 256) {
        float2 acc = make_float2(0.f, 0.f);
        if (q > 0) {
            const int hi = min(nb, t_r - q);  // R frames q + b must exist: b < t_r - q
            int b = 0;
            for (; b + 4 <= hi; b += 4) {
                float2 d[4], r[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    d[j] = __ldg(D + (size_t)(b + j + 1) * bins + k);
                    r[j] = __ldg(R + (size_t)(q + b + j) * bins + k);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {  // conj(d) * r
                    acc.x = fmaf(d[j].x, r[j].x, fmaf(d[j].y, r[j].y, acc.x));
                    acc.y = fmaf(d[j].x, r[j].y, fmaf(-d[j].y, r[j].x, acc.y));
                }
            }
            for (; b < hi; ++b) {
                const float2 d = __ldg(D + (size_t)(b + 1) * bins + k), r = __ldg(R + (size_t)(q + b) * bins + k);
                acc.x = fmaf(d.x, r.x, fmaf(d.y, r.y, acc.x));
                acc.y = fmaf(d.x, r.y, fmaf(-d.y, r.x, acc.y));
            }
            if (k & 1) { acc.x = -acc.x; acc.y = -acc.y; }
        }
        Z[(size_t)q * bins + k] = acc;
    }
}

// first maximum of x[0 .. n): *out = argmax - sub (single CTA; n is a few hundred thousand at most)
__global__ void __launch_bounds__(1024)
argmax_first_kernel(const float* __restrict__ x, int n, int sub, int* __restrict__ out) {
    __shared__ float rv[32];
    __shared__ int ri[32];
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int k = threadIdx.x; k < n; k += 1024) {
        const float v = __ldg(x + k);
        if (v > bv) { bv = v; bi = k; }  // (k increases per thread: the first maximum is kept)
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { rv[threadIdx.x >> 5] = bv; ri[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 32; ++k)
            if (rv[k] > bv || (rv[k] == bv && ri[k] < bi)) { bv = rv[k]; bi = ri[k]; }
        *out = bi - sub;
    }
}

}  // namespace adv

using namespace adv;

extern "C" {

int adv_row_stats_parts(int n) { return n <= 0 ? 0 : (n + kRowChunk - 1) / kRowChunk; }

int adv_row_stats(const float* in, int batch, int n, double* stats, void* stream) {
    if (!in || !stats || batch <= 0 || n <= 1) return ADV_ERR_INVALID;
    const int parts = adv_row_stats_parts(n);
    if (n % 4 == 0 && reinterpret_cast<uintptr_t>(in) % 16 == 0)
        row_stats4_kernel<<<dim3(parts, batch), kPwThreads, 0, (cudaStream_t)stream>>>(in, n, parts, stats);
    else
        row_stats_kernel<<<dim3(parts, batch), kPwThreads, 0, (cudaStream_t)stream>>>(in, n, parts, stats);
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

int adv_normalize(const float* in, float* out, int batch, int n, const double* stats, int parts, int width,
                  int col, void* stream) {
    if (!in || !out || !stats || batch <= 0 || n <= 1 || parts <= 0 || width < 2 || col < 0 || col + 2 > width)
        return ADV_ERR_INVALID;
    const int chunks = (n + kRowChunk - 1) / kRowChunk;
    ADV_CUDA_CHECK(launch_pdl(normalize_kernel<kPwThreads>, dim3(chunks, batch, 1), dim3(kPwThreads), 0, (cudaStream_t)stream,
                              in, out, in, out, n, stats, parts, width, col, LmacJob{nullptr, nullptr, nullptr, 0, 0, nullptr, nullptr}));
    return ADV_OK;
}

int adv_normalize_pair(float* rel, float* irr, int batch, int n, const double* stats, int parts, void* stream) {
    if (!rel || !irr || !stats || batch <= 0 || n <= 1 || parts <= 0) return ADV_ERR_INVALID;
    const int chunks = (n + kRowChunk - 1) / kRowChunk;
    // 256-thread CTAs by default; ADV_NORM_THREADS=128 selects 128-thread CTAs (one fits next to an explain kernel
    // capped at 120 registers - measured: the cap costs the explain kernel more than the overlap returns)
    static const char* nt_env = ADV_AB_ENV("ADV_NORM_THREADS");
    if (nt_env && nt_env[0] == '1')
        ADV_CUDA_CHECK(launch_pdl(normalize_kernel<128>, dim3(chunks, batch, 2), dim3(128), 0, (cudaStream_t)stream, rel, rel,
                                  irr, irr, n, stats, parts, 4, 0, LmacJob{nullptr, nullptr, nullptr, 0, 0, nullptr, nullptr}));
    else
        ADV_CUDA_CHECK(launch_pdl(normalize_kernel<256>, dim3(chunks, batch, 2), dim3(256), 0, (cudaStream_t)stream, rel, rel,
                                  irr, irr, n, stats, parts, 4, 0, LmacJob{nullptr, nullptr, nullptr, 0, 0, nullptr, nullptr}));
    return ADV_OK;
}

int adv_normalize_pair_lmac(float* rel, float* irr, int batch, int n, const double* stats, int parts, const float* p,
                            const float* theta, const float* q, int n_logits, int flags, float* scores, double* sums,
                            void* stream) {
    if (!rel || !irr || !stats || batch <= 0 || n <= 1 || parts <= 0 || !p || !theta || !q || !sums || n_logits <= 0)
        return ADV_ERR_INVALID;
    if (n_logits > kLmacPerBlock) return ADV_ERR_UNSUPPORTED;   // more than one metric CTA: use adv_lmac_reduce
    const int chunks = (n + kRowChunk - 1) / kRowChunk;
    const LmacJob job{p, theta, q, n_logits, flags, scores, sums};
    ADV_CUDA_CHECK(launch_pdl(normalize_kernel<256>, dim3(chunks + 1, batch, 2), dim3(256), 0, (cudaStream_t)stream, rel, rel,
                              irr, irr, n, stats, parts, 4, 0, job));
    return ADV_OK;
}

int adv_lmac_blocks(int n) { return n <= 0 ? 0 : (n + kLmacPerBlock - 1) / kLmacPerBlock; }

int adv_lmac_reduce(const float* p, const float* theta, const float* q, int n, int flags, float* scores,
                    double* sums, double* block_partials, unsigned int* counter, void* stream) {
    if (!p || !theta || !q || !sums || !block_partials || !counter || n <= 0) return ADV_ERR_INVALID;
    ADV_CUDA_CHECK(launch_pdl(lmac_kernel, dim3(adv_lmac_blocks(n)), dim3(kPwThreads), 0, (cudaStream_t)stream, p, theta, q,
                              n, flags, scores, sums, block_partials, counter));
    return ADV_OK;
}

int adv_mask_apply(const float* mag, const float* phase, const float* mask, int batch, int T, int F, int Fm, int Tm,
                   int mode, adv_c64* rel, adv_c64* irr, void* stream) {
    if (!mag || !phase || !mask || !rel || !irr || batch <= 0 || T <= 0 || F <= 0) return ADV_ERR_INVALID;
    if (Fm > F || Tm > T || Fm <= 0 || Tm <= 0) return ADV_ERR_SHAPE;
    dim3 grid((F + 31) / 32, (T + 31) / 32, batch);
    if (mode == ADV_MASK_LOG1P)
        mask_apply_kernel<ADV_MASK_LOG1P><<<grid, kPwThreads, 0, (cudaStream_t)stream>>>(
            mag, phase, mask, T, F, Fm, Tm, (float2*)rel, (float2*)irr);
    else if (mode == ADV_MASK_LINEAR)
        mask_apply_kernel<ADV_MASK_LINEAR><<<grid, kPwThreads, 0, (cudaStream_t)stream>>>(
            mag, phase, mask, T, F, Fm, Tm, (float2*)rel, (float2*)irr);
    else
        return ADV_ERR_INVALID;
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

int adv_td_mask(const float* wave, const float* attr, int batch, int n, float* mask_out, float* rel, float* irr,
                float* rowmax, void* stream) {
    if (!wave || !attr || !rel || !irr || !rowmax || batch <= 0 || n <= 0) return ADV_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    ADV_CUDA_CHECK(cudaMemsetAsync(rowmax, 0, sizeof(float) * batch, s));
    const int chunks = (n + kRowChunk - 1) / kRowChunk;
    const uintptr_t al = reinterpret_cast<uintptr_t>(wave) | reinterpret_cast<uintptr_t>(attr) | reinterpret_cast<uintptr_t>(rel) |
                         reinterpret_cast<uintptr_t>(irr) | reinterpret_cast<uintptr_t>(mask_out);
    static const bool scalar = ADV_AB_ENV("ADV_TD_MASK_SCALAR") != nullptr;
    if (!scalar && n % 4 == 0 && al % 16 == 0) {
        rowmax_abs4_kernel<<<dim3(chunks, batch), kPwThreads, 0, s>>>(attr, n, rowmax);
        td_mask4_kernel<<<dim3(chunks, batch), kPwThreads, 0, s>>>(wave, attr, n, rowmax, mask_out, rel, irr);
    } else {
        rowmax_abs_kernel<<<dim3(chunks, batch), kPwThreads, 0, s>>>(attr, n, rowmax);
        td_mask_kernel<<<dim3(chunks, batch), kPwThreads, 0, s>>>(wave, attr, n, rowmax, mask_out, rel, irr);
    }
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

int adv_mask_head(const float* y1, const float* w, const float* bias, int batch, int channels, int64_t hw, float* mask,
                  void* stream) {
    if (!y1 || !w || !bias || !mask || batch <= 0 || channels <= 0 || hw <= 0) return ADV_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec = (hw % 4 == 0) && ((reinterpret_cast<uintptr_t>(y1) | reinterpret_cast<uintptr_t>(mask)) % 16 == 0);
    const int64_t work = vec ? hw / 4 : hw;
    int gx = (int)((work + kPwThreads - 1) / kPwThreads);
    if (gx > 148 * 16) gx = 148 * 16;
    if (channels == 32) {
        if (vec) mask_head_kernel<32, true><<<dim3(gx, batch), kPwThreads, 0, s>>>(y1, w, bias, hw, mask);
        else mask_head_kernel<32, false><<<dim3(gx, batch), kPwThreads, 0, s>>>(y1, w, bias, hw, mask);
    } else {
        mask_head_generic_kernel<<<dim3(gx, batch), kPwThreads, 0, s>>>(y1, w, bias, channels, hw, mask);
    }
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

int adv_band_swap(const adv_c64* real, const adv_c64* voc, int batch, int T, int F, int f_lo, int f_hi, adv_c64* out,
                  void* stream) {
    if (!real || !voc || !out || batch <= 0 || T <= 0 || F <= 0) return ADV_ERR_INVALID;
    const int64_t total = (int64_t)batch * T * F;
    int gx = (int)((total + kPwThreads - 1) / kPwThreads);
    if (gx > 148 * 16) gx = 148 * 16;
    band_swap_kernel<<<gx, kPwThreads, 0, (cudaStream_t)stream>>>((const float2*)real, (const float2*)voc, total, F,
                                                                 f_lo, f_hi, (float2*)out);
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

}  // extern "C"

extern "C" int adv_xcorr_blocks(int n_ref, int n_deg) {
    if (n_ref <= 0 || n_deg <= 0) return 0;
    return (n_ref + n_deg + 1 + kXcLags - 1) / kXcLags;
}

extern "C" int adv_xcorr_shift(const float* ref, int n_ref, const float* deg, int n_deg, float* ws_val, int* ws_idx, int* shift,
                    void* stream) {
    if (!ref || !deg || !ws_val || !ws_idx || !shift || n_ref <= 0 || n_deg <= 0) return ADV_ERR_INVALID;
    const int blocks = adv_xcorr_blocks(n_ref, n_deg);
    xcorr_partial_kernel<<<blocks, kXcThreads, 0, (cudaStream_t)stream>>>(ref, n_ref, deg, n_deg, ws_val, ws_idx);
    xcorr_final_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(ws_val, ws_idx, blocks, n_deg, shift);
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

extern "C" int adv_xcorr_fd_mac(const adv_c64* D, int nb, const adv_c64* R, int t_r, adv_c64* Z, int nq, int bins,
                               void* stream) {
    if (!D || !R || !Z || nb <= 0 || t_r <= 0 || nq <= 0 || bins <= 0) return ADV_ERR_INVALID;
    xcorr_fd_mac_kernel<<<nq + 1, 256, 0, (cudaStream_t)stream>>>((const float2*)D, nb, (const float2*)R, t_r, (float2*)Z,
                                                                  bins);
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

extern "C" int adv_argmax_first(const float* x, int n, int sub, int* out, void* stream) {
    if (!x || !out || n <= 0) return ADV_ERR_INVALID;
    argmax_first_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(x, n, sub, out);
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

extern "C" int adv_mask_grad_linear(const adv_c64* X, int64_t sb, int64_t st, int64_t sf, const adv_c64* A, int batch, int F,
                                    int T, int Fm, int Tm, float* gm, void* stream) {
    if (!X || !A || !gm || batch <= 0 || F <= 0 || T <= 0 || Fm <= 0 || Tm <= 0 || Fm > F || Tm > T) return ADV_ERR_INVALID;
    dim3 grid((Fm + 31) / 32, (Tm + 31) / 32, batch);
    mask_grad_linear_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float2*)X, sb, st, sf, (const float2*)A, F, T, Fm,
                                                                  Tm, gm);
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

extern "C" int adv_band_swap_multi(const adv_c64* real, const adv_c64* voc, int batch, int T, int F, const int* edges,
                                   int n_bands, adv_c64* out, void* stream) {
    if (!real || !voc || !out || !edges || batch <= 0 || T <= 0 || F <= 0 || n_bands <= 0) return ADV_ERR_INVALID;
    const int64_t total = (int64_t)batch * T * F;
    int gx = (int)((total + kPwThreads - 1) / kPwThreads);
    if (gx > 148 * 16) gx = 148 * 16;
    band_swap_multi_kernel<<<gx, kPwThreads, 0, (cudaStream_t)stream>>>((const float2*)real, (const float2*)voc, total, F,
                                                                       edges, n_bands, (float2*)out);
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// AudioProcessor.load_audio, batched (audioprocessor.py:49-63): PCM16 -> float (/ 32768, torchaudio.load's normalisation)
// -> optional sinc resampling (torchaudio.transforms.Resample defaults: sinc_interp_hann, lowpass_filter_width 6, rolloff
// 0.99) -> zero-pad / crop to n_out samples, for B ragged clips in one launch.  The polyphase filter h[new][K]
// (K = 2 width + orig) is built on the host with torchaudio's own formula (audioprocessor.resample_filter);
// out[b][j] = sum_k h[j % new][k] * x_b[(j / new) * orig + k - width]   for j < ceil(new * len_b / orig), zero beyond;
// only the taps inside [range[p].x, range[p].y) are non-zero for phase p (the formula clamps t to the window's support:
// ~34 of 475 taps for 44.1 kHz -> 16 kHz) and only those are summed.
// ---------------------------------------------------------------------------------------------------------------------
template <class TIn>
__device__ __forceinline__ float sample_as_float(const TIn* p, long i);
template <>
__device__ __forceinline__ float sample_as_float<short>(const short* p, long i) { return (float)p[i] * (1.0f / 32768.0f); }
template <>
__device__ __forceinline__ float sample_as_float<float>(const float* p, long i) { return p[i]; }

template <class TIn>
__global__ void resample_rows_kernel(const TIn* __restrict__ in, const long long* __restrict__ offs, const int* __restrict__ lens,
                                     int orig, int newf, int width, int K, const float* __restrict__ h,
                                     const int2* __restrict__ range, int n_out, float* __restrict__ out) {
    const int b = blockIdx.y;
    const int len = lens[b];
    const TIn* x = in + offs[b];
    const long long target = orig == newf ? (long long)len : ((long long)newf * len + orig - 1) / orig;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_out; j += gridDim.x * blockDim.x) {
        float acc = 0.0f;
        if (j < target) {
            if (orig == newf) {
                acc = sample_as_float<TIn>(x, j);
            } else {
                const int n = j / newf, p = j - n * newf;
                const int2 r = range[p];
                const long base = (long)n * orig - width;
                const float* hp = h + (size_t)p * K;
                for (int k = r.x; k < r.y; ++k) {
                    const long i = base + k;
                    if (i >= 0 && i < len) acc = fmaf(hp[k], sample_as_float<TIn>(x, i), acc);
                }
            }
        }
        out[(size_t)b * n_out + j] = acc;
    }
}

extern "C" int adv_resample_rows(const void* in, int in_is_pcm16, const long long* offs, const int* lens, int batch, int orig,
                                 int newf, int width, const float* h, const int* range, int n_out, float* out, void* stream) {
    if (!in || !offs || !lens || !out || batch <= 0 || n_out <= 0 || orig <= 0 || newf <= 0) return ADV_ERR_INVALID;
    if (orig != newf && (!h || !range || width <= 0)) return ADV_ERR_INVALID;
    const int K = 2 * width + orig;
    int gx = (n_out + kPwThreads - 1) / kPwThreads;
    if (gx > 148 * 8) gx = 148 * 8;
    dim3 grid(gx, batch);
    if (in_is_pcm16)
        resample_rows_kernel<short><<<grid, kPwThreads, 0, (cudaStream_t)stream>>>((const short*)in, offs, lens, orig, newf, width,
                                                                                   K, h, (const int2*)range, n_out, out);
    else
        resample_rows_kernel<float><<<grid, kPwThreads, 0, (cudaStream_t)stream>>>((const float*)in, offs, lens, orig, newf, width,
                                                                                   K, h, (const int2*)range, n_out, out);
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}
