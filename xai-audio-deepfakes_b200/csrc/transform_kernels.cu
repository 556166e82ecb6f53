// STFT, iSTFT (overlap-add) and the fused explain kernel for sm_100a.
//
// Reference call sites these kernels replace:
//   stft_kernel    - torch.stft + .abs() + .angle()      audioprocessor.py:102-110
//   istft_kernel   - torch.istft                         audioprocessor.py:123-129
//   explain_kernel - STFT -> mask/(1-mask) apply -> 2x iSTFT, LMAC_metrics.py:136-157 (log1p mode),
//                    loss_function.py:36-47 (linear mode)
//
// Data layout in HBM: waveforms [B][n] fp32 rows; spectra frame-major [B][T][F] complex64 (what
// torch.stft itself produces); masks [B][Fm][Tm] fp32 as the mask network emits them.
//
// CTA = 256 threads = 16 units (n_fft 512) or 8 units (n_fft 1024); a unit is the lane group that
// cooperates on one complex FFT (fft_core.cuh).  Two real frames ride in one complex FFT; in the
// explain kernel the masked-in and masked-out spectra of ONE frame ride in one inverse FFT.
// Overlap-add is output-stationary: a CTA owns a span of output samples, recomputes the few halo
// frames that overlap it, and adds frames into a shared-memory accumulator in `phases` rounds
// (frames t, t+phases, ... never overlap) - no global or shared atomics, deterministic order.
#include <mutex>
#include <unordered_map>
#include "adv_internal.cuh"
#include "fft_core.cuh"

namespace adv {

__device__ __forceinline__ int floordiv(int a, int b) {  // b > 0
    int q = a / b;
    return (a % b != 0 && a < 0) ? q - 1 : q;
}
__device__ __forceinline__ int ceildiv(int a, int b) { return floordiv(a + b - 1, b); }

template <int LANES>
struct TwSmem {
    const float2* p;
    int l;
    __device__ __forceinline__ float2 operator()(int k1) const { return p[k1 * LANES + l]; }
};

template <int NF>
__device__ __forceinline__ void split_regs(const float2* v, int l, float2* xa, float2* xb) {
    if constexpr (NF == 512) {
        split512(v, l, xa, xb);
    } else {
        float2 send[16], recv[16];
        split1024_pre(v, send);
        const int src = (32 - l) & 31;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            recv[i].x = __shfl_sync(0xffffffffu, send[i].x, src);
            recv[i].y = __shfl_sync(0xffffffffu, send[i].y, src);
        }
        split1024_post(v, l, recv, xa, xb);
    }
}
template <int NF>
__device__ __forceinline__ void merge_regs(float2* v, int l, const float2* ya, const float2* yb) {
    if constexpr (NF == 512) {
        merge512(v, l, ya, yb);
    } else {
        float2 send[16], recv[16];
        merge1024_pre(v, l, ya, yb, send);
        const int src = (32 - l) & 31;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            recv[i].x = __shfl_sync(0xffffffffu, send[i].x, src);
            recv[i].y = __shfl_sync(0xffffffffu, send[i].y, src);
        }
        merge1024_post(v, l, recv);
    }
}

// Stage `seglen` waveform samples starting at original index `base` (may be negative / past the end:
// torch.stft's centre=True reflect padding) into shared memory.
__device__ __forceinline__ void load_segment(float* seg, int seglen, const float* __restrict__ row, int base,
                                             int n_in) {
    for (int i = threadIdx.x; i < seglen; i += kThreads) {
        int idx = base + i;
        if (idx < 0) idx = -idx;
        else if (idx >= n_in) idx = 2 * (n_in - 1) - idx;
        seg[i] = (idx >= 0 && idx < n_in) ? __ldg(row + idx) : 0.0f;
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Sum NQ doubles per thread over the CTA; result valid in thread 0.  red: >= NQ * 8 doubles of smem.
template <int NQ>
__device__ __forceinline__ void block_sum(double (&q)[NQ], double* red) {
    const int w = threadIdx.x >> 5, ln = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
        q[i] = warp_sum(q[i]);
        if (ln == 0) red[i * (kThreads / 32) + w] = q[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
            double s = 0.0;
            for (int k = 0; k < kThreads / 32; ++k) s += red[i * (kThreads / 32) + k];
            q[i] = s;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// STFT
// ------------------------------------------------------------------------------------------------
template <int NF>
struct StftSmem {
    using G = Geo<NF>;
    static constexpr int UNITS = kThreads / G::LANES;
    static constexpr int FT = 2 * UNITS;  // frames per CTA
    static size_t bytes(int hop) {
        return sizeof(float2) * (32 * G::LANES + UNITS * G::SCRATCH) + sizeof(float) * (NF + (FT - 1) * hop + NF);
    }
};

template <int NF, bool MAG, bool PHASE>
__global__ void __launch_bounds__(kThreads)
stft_kernel(PlanDev P, const float* __restrict__ wav, int64_t wav_stride, float2* __restrict__ X,
            float* __restrict__ mag, float* __restrict__ phase) {
    using G = Geo<NF>;
    constexpr int UNITS = StftSmem<NF>::UNITS, FT = StftSmem<NF>::FT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* tw_s = reinterpret_cast<float2*>(smem_raw);
    float2* scratch = tw_s + 32 * G::LANES;
    float* win_s = reinterpret_cast<float*>(scratch + UNITS * G::SCRATCH);
    float* seg = win_s + NF;

    const int tid = threadIdx.x, b = blockIdx.y, t0 = blockIdx.x * FT;
    for (int i = tid; i < 32 * G::LANES; i += kThreads) tw_s[i] = P.tw[i];
    for (int i = tid; i < NF; i += kThreads) win_s[i] = P.window[i];
    load_segment(seg, (FT - 1) * P.hop + NF, wav + (size_t)b * wav_stride, t0 * P.hop - NF / 2, P.n_in);
    __syncthreads();

    const int u = tid / G::LANES, l = tid % G::LANES;
    float2* my = scratch + u * G::SCRATCH;
    const int fa = t0 + 2 * u, fb = fa + 1;  // adjacent frames share one complex FFT
    const float* sa = seg + (2 * u) * P.hop;
    const float* sb = sa + P.hop;
    float2 v[32];
#pragma unroll
    for (int n1 = 0; n1 < 32; ++n1) {
        const int n = n1 * G::R2 + l;
        const float w = win_s[n];
        v[n1] = make_float2(sa[n] * w, sb[n] * w);
    }
    fwd_phase_a<NF>(v, l, TwSmem<G::LANES>{tw_s, l}, my);
    __syncwarp();
    fwd_phase_b<NF>(v, l, my);
    float2 xa[17], xb[17];
    split_regs<NF>(v, l, xa, xb);

    constexpr int F = G::NBINS;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int t = half ? fb : fa;
        if (t >= P.T) continue;
        const float2* x = half ? xb : xa;
        const size_t row = ((size_t)b * P.T + t) * F;
#pragma unroll
        for (int i = 0; i < 17; ++i) {
            const int bin = bin_of<NF>(l, i);
            if (bin < 0) continue;
            X[row + bin] = x[i];
            if (MAG) mag[row + bin] = sqrtf(x[i].x * x[i].x + x[i].y * x[i].y);
            if (PHASE) phase[row + bin] = atan2f(x[i].y, x[i].x);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// shared pieces of the overlap-add kernels
// ------------------------------------------------------------------------------------------------
struct TileGeom {
    int s0, s1;      // output samples [s0, s1)
    int p0;          // padded position of s0
    int t_lo, t_hi;  // frames overlapping the tile (inclusive)
};
template <int NF>
__device__ __forceinline__ TileGeom tile_geom(const PlanDev& P, const Tiling& TL, int tile) {
    TileGeom g;
    const int S = TL.hops_per_tile * P.hop;
    g.s0 = tile * S;
    g.s1 = min(g.s0 + S, P.n_out);
    g.p0 = g.s0 + NF / 2;
    const int p1 = g.s1 + NF / 2;
    g.t_lo = max(0, ceildiv(g.p0 - P.whi + 1, P.hop));
    g.t_hi = min(P.T - 1, floordiv(p1 - 1 - P.wlo, P.hop));
    return g;
}

// ------------------------------------------------------------------------------------------------
// iSTFT
// ------------------------------------------------------------------------------------------------
template <int NF>
struct IstftSmem {
    using G = Geo<NF>;
    static constexpr int UNITS = kThreads / G::LANES;
    static constexpr int FT = 2 * UNITS;
    static size_t bytes(int tile_samples) {
        return sizeof(float2) * (32 * G::LANES + UNITS * G::SCRATCH) + sizeof(float) * (NF + tile_samples) +
               sizeof(double) * 2 * (kThreads / 32);
    }
};

template <int NF>
__global__ void __launch_bounds__(kThreads)
istft_kernel(PlanDev P, Tiling TL, const float2* __restrict__ X, int64_t sb, int64_t st, int64_t sf,
             float* __restrict__ out, double* __restrict__ stats) {
    using G = Geo<NF>;
    constexpr int UNITS = IstftSmem<NF>::UNITS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* tw_s = reinterpret_cast<float2*>(smem_raw);
    float2* scratch = tw_s + 32 * G::LANES;
    double* red = reinterpret_cast<double*>(scratch + UNITS * G::SCRATCH);
    float* win_s = reinterpret_cast<float*>(red + 2 * (kThreads / 32));
    float* ola = win_s + NF;

    const int tid = threadIdx.x, b = blockIdx.y;
    const TileGeom g = tile_geom<NF>(P, TL, blockIdx.x);
    const int S = g.s1 - g.s0;
    for (int i = tid; i < 32 * G::LANES; i += kThreads) tw_s[i] = P.tw[i];
    for (int i = tid; i < NF; i += kThreads) win_s[i] = P.window[i];
    for (int i = tid; i < S; i += kThreads) ola[i] = 0.0f;

    const int u = tid / G::LANES, l = tid % G::LANES;
    float2* my = scratch + u * G::SCRATCH;
    const int fa = g.t_lo + u, fb = fa + UNITS;
    const bool va = fa <= g.t_hi, vb = fb <= g.t_hi;

    float2 ya[17], yb[17];
    {
        const float2* xa_p = X + (size_t)b * sb + (size_t)fa * st;
        const float2* xb_p = X + (size_t)b * sb + (size_t)fb * st;
#pragma unroll
        for (int i = 0; i < 17; ++i) {
            const int bin = bin_of<NF>(l, i);
            const float2 z = make_float2(0.f, 0.f);
            ya[i] = (va && bin >= 0) ? __ldg(xa_p + (size_t)bin * sf) : z;
            yb[i] = (vb && bin >= 0) ? __ldg(xb_p + (size_t)bin * sf) : z;
        }
    }
    __syncthreads();  // tables + zeroed accumulator visible

    float2 v[32];
    merge_regs<NF>(v, l, ya, yb);
    inv_phase_a<NF>(v, l, my);
    __syncwarp();
    inv_phase_b<NF>(v, l, TwSmem<G::LANES>{tw_s, l}, my);

    // overlap-add, `phases` rounds per packed frame; frames a and b = a + UNITS share a round when
    // UNITS is a multiple of `phases`
    const bool same = (UNITS % P.phases) == 0;
    for (int pass = 0; pass < (same ? 1 : 2); ++pass) {
        for (int ph = 0; ph < P.phases; ++ph) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                if (!same && half != pass) continue;
                const int t = half ? fb : fa;
                const bool ok = (half ? vb : va) && (t % P.phases) == ph;
                if (ok) {
                    const int off = t * P.hop - g.p0;
#pragma unroll
                    for (int n1 = 0; n1 < 32; ++n1) {
                        const int n = n1 * G::R2 + l;
                        const int q = off + n;
                        if (n >= P.wlo && n < P.whi && q >= 0 && q < S)
                            ola[q] += (half ? v[n1].y : v[n1].x) * win_s[n];
                    }
                }
            }
            __syncthreads();
        }
    }

    double acc[2] = {0.0, 0.0};
    float* orow = out + (size_t)b * P.n_out + g.s0;
    const float* env = P.inv_env + g.s0;
    for (int q = tid; q < S; q += kThreads) {
        const float y = ola[q] * __ldg(env + q);
        orow[q] = y;
        acc[0] += (double)y;
        acc[1] += (double)y * (double)y;
    }
    if (stats != nullptr) {
        block_sum<2>(acc, red);
        if (tid == 0) {
            double* srow = stats + ((size_t)b * TL.tiles + blockIdx.x) * 2;
            srow[0] = acc[0];
            srow[1] = acc[1];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// fused explain: wave (or spectrum) + mask -> masked-in / masked-out waveforms
// ------------------------------------------------------------------------------------------------
template <int NF>
struct ExplainSmem {
    using G = Geo<NF>;
    static constexpr int UNITS = kThreads / G::LANES;
    static constexpr int FT = 2 * UNITS;
    static constexpr int MP = FT + 1;  // mask tile pitch (odd)
    static size_t bytes(int hop, int tile_samples, bool from_spec) {
        size_t b = sizeof(float2) * (32 * G::LANES + UNITS * G::SCRATCH + tile_samples) +
                   sizeof(double) * 4 * (kThreads / 32) + sizeof(float) * (NF + G::NBINS * MP);
        if (!from_spec) b += sizeof(float) * ((FT - 1) * hop + NF);
        return b;
    }
};

template <int MODE>
__device__ __forceinline__ void mask_gains(float2 x, float m, float& g_rel, float& g_irr) {
    if (MODE == ADV_MASK_LINEAR) {  // loss_function.py:38-45: m*|X|*e^{i phi} = m*X
        g_rel = m;
        g_irr = 1.0f - m;
        return;
    }
    // LMAC_metrics.py:138-143: expm1(m*log1p(a)) * e^{i phi} = X * expm1(m*log1p(a)) / a
    const float a = sqrtf(x.x * x.x + x.y * x.y);
    const float lm = log1pf(a);
    const float er = expm1f(m * lm), ei = expm1f((1.0f - m) * lm);
    if (a > 1e-30f) {
        const float ia = 1.0f / a;
        g_rel = er * ia;
        g_irr = ei * ia;
    } else {  // limit a -> 0: expm1(m*log1p(a))/a -> m (the product with X is 0 either way)
        g_rel = m;
        g_irr = 1.0f - m;
    }
}

template <int NF, int MODE, bool FROM_SPEC>
__global__ void __launch_bounds__(kThreads, 1)
explain_kernel(PlanDev P, Tiling TL, const float* __restrict__ wav, int64_t wav_stride,
               const float2* __restrict__ X, int64_t sb, int64_t st, int64_t sf,
               const float* __restrict__ mask, int Fm, int Tm,
               float* __restrict__ rel, float* __restrict__ irr, double* __restrict__ stats) {
    using G = Geo<NF>;
    using SM = ExplainSmem<NF>;
    constexpr int UNITS = SM::UNITS, FT = SM::FT, MP = SM::MP, F = G::NBINS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* tw_s = reinterpret_cast<float2*>(smem_raw);
    float2* scratch = tw_s + 32 * G::LANES;
    float2* ola = scratch + UNITS * G::SCRATCH;  // .x masked-in, .y masked-out
    const int Smax = TL.hops_per_tile * P.hop;
    double* red = reinterpret_cast<double*>(ola + Smax);
    float* win_s = reinterpret_cast<float*>(red + 4 * (kThreads / 32));
    float* mask_s = win_s + NF;
    float* seg = mask_s + F * MP;

    const int tid = threadIdx.x, b = blockIdx.y;
    const TileGeom g = tile_geom<NF>(P, TL, blockIdx.x);
    const int S = g.s1 - g.s0;
    for (int i = tid; i < 32 * G::LANES; i += kThreads) tw_s[i] = P.tw[i];
    for (int i = tid; i < NF; i += kThreads) win_s[i] = P.window[i];
    for (int i = tid; i < S; i += kThreads) ola[i] = make_float2(0.f, 0.f);
    {   // mask tile [F][FT] <- mask[b][f][t_lo + c], zero outside the mask / clip
        const float* mrow = mask + (size_t)b * Fm * Tm;
        for (int e = tid; e < F * FT; e += kThreads) {
            const int f = e / FT, c = e % FT;
            const int t = g.t_lo + c;
            mask_s[f * MP + c] = (f < Fm && t < Tm && t <= g.t_hi) ? __ldg(mrow + (size_t)f * Tm + t) : 0.0f;
        }
    }
    if (!FROM_SPEC)
        load_segment(seg, (FT - 1) * P.hop + NF, wav + (size_t)b * wav_stride, g.t_lo * P.hop - NF / 2, P.n_in);
    __syncthreads();

    const int u = tid / G::LANES, l = tid % G::LANES;
    float2* my = scratch + u * G::SCRATCH;
    const int fa = g.t_lo + u, fb = fa + UNITS;
    const bool va = fa <= g.t_hi, vb = fb <= g.t_hi;

    float2 v[32];
    float2 xa[17], xb[17];
    if (FROM_SPEC) {
        const float2* xa_p = X + (size_t)b * sb + (size_t)fa * st;
        const float2* xb_p = X + (size_t)b * sb + (size_t)fb * st;
#pragma unroll
        for (int i = 0; i < 17; ++i) {
            const int bin = bin_of<NF>(l, i);
            const float2 z = make_float2(0.f, 0.f);
            xa[i] = (va && bin >= 0) ? __ldg(xa_p + (size_t)bin * sf) : z;
            xb[i] = (vb && bin >= 0) ? __ldg(xb_p + (size_t)bin * sf) : z;
        }
    } else {
        const float* sa = seg + u * P.hop;
        const float* sbp = sa + UNITS * P.hop;
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
            const int n = n1 * G::R2 + l;
            const float w = win_s[n];
            v[n1] = make_float2(va ? sa[n] * w : 0.f, vb ? sbp[n] * w : 0.f);
        }
        fwd_phase_a<NF>(v, l, TwSmem<G::LANES>{tw_s, l}, my);
        __syncwarp();
        fwd_phase_b<NF>(v, l, my);
        __syncwarp();
        split_regs<NF>(v, l, xa, xb);
    }

#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int t = half ? fb : fa;
        const bool valid = half ? vb : va;
        const float2* x = half ? xb : xa;
        const int col = t - g.t_lo;
        float2 yr[17], yi[17];
#pragma unroll
        for (int i = 0; i < 17; ++i) {
            const int bin = bin_of<NF>(l, i);
            const float m = (bin >= 0 && valid) ? mask_s[bin * MP + col] : 0.0f;
            float gr, gi;
            mask_gains<MODE>(x[i], m, gr, gi);
            const bool live = bin >= 0;
            yr[i] = live ? make_float2(x[i].x * gr, x[i].y * gr) : make_float2(0.f, 0.f);
            yi[i] = live ? make_float2(x[i].x * gi, x[i].y * gi) : make_float2(0.f, 0.f);
        }
        merge_regs<NF>(v, l, yr, yi);
        inv_phase_a<NF>(v, l, my);
        __syncwarp();
        inv_phase_b<NF>(v, l, TwSmem<G::LANES>{tw_s, l}, my);
        // v[n1] = n_fft * (rel[n], irr[n]) of frame t
        for (int ph = 0; ph < P.phases; ++ph) {
            if (valid && (t % P.phases) == ph) {
                const int off = t * P.hop - g.p0;
#pragma unroll
                for (int n1 = 0; n1 < 32; ++n1) {
                    const int n = n1 * G::R2 + l;
                    const int q = off + n;
                    if (n >= P.wlo && n < P.whi && q >= 0 && q < S) {
                        const float w = win_s[n];
                        float2 o = ola[q];
                        o.x += v[n1].x * w;
                        o.y += v[n1].y * w;
                        ola[q] = o;
                    }
                }
            }
            __syncthreads();
        }
    }

    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    float* rrow = rel + (size_t)b * P.n_out + g.s0;
    float* irow = irr + (size_t)b * P.n_out + g.s0;
    const float* env = P.inv_env + g.s0;
    for (int q = tid; q < S; q += kThreads) {
        const float e = __ldg(env + q);
        const float2 o = ola[q];
        const float yr = o.x * e, yi = o.y * e;
        rrow[q] = yr;
        irow[q] = yi;
        acc[0] += (double)yr;
        acc[1] += (double)yr * (double)yr;
        acc[2] += (double)yi;
        acc[3] += (double)yi * (double)yi;
    }
    if (stats != nullptr) {
        block_sum<4>(acc, red);
        if (tid == 0) {
            double* srow = stats + ((size_t)b * TL.tiles + blockIdx.x) * 4;
            srow[0] = acc[0];
            srow[1] = acc[1];
            srow[2] = acc[2];
            srow[3] = acc[3];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------
// cudaFuncSetAttribute once per (kernel, size high-water mark): keeps the launch path free of
// attribute calls in steady state (and inside CUDA-graph capture)
template <class K>
static int set_smem(K kernel, size_t bytes) {
    static std::mutex mu;
    static std::unordered_map<const void*, size_t> high;
    if (bytes > 227 * 1024) return ADV_ERR_UNSUPPORTED;
    std::lock_guard<std::mutex> lock(mu);
    size_t& cur = high[reinterpret_cast<const void*>(kernel)];
    if (bytes > cur) {
        ADV_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        cur = bytes;
    }
    return ADV_OK;
}

template <int NF>
static int launch_stft_nf(const adv_plan* p, const float* wav, int64_t wav_stride, int batch, float2* X,
                          float* mag, float* phase, cudaStream_t s) {
    constexpr int FT = StftSmem<NF>::FT;
    const size_t smem = StftSmem<NF>::bytes(p->d.hop);
    dim3 grid((p->d.T + FT - 1) / FT, batch);
    int rc;
#define ADV_LAUNCH_STFT(M, PH)                                                           \
    do {                                                                                 \
        if ((rc = set_smem(stft_kernel<NF, M, PH>, smem)) != ADV_OK) return rc;          \
        stft_kernel<NF, M, PH><<<grid, kThreads, smem, s>>>(p->d, wav, wav_stride, X, mag, phase); \
    } while (0)
    if (mag && phase) ADV_LAUNCH_STFT(true, true);
    else if (mag) ADV_LAUNCH_STFT(true, false);
    else if (phase) ADV_LAUNCH_STFT(false, true);
    else ADV_LAUNCH_STFT(false, false);
#undef ADV_LAUNCH_STFT
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

int launch_stft(const adv_plan* p, const float* wav, int64_t wav_stride, int batch, float2* X, float* mag,
                float* phase, cudaStream_t s) {
    return p->d.n_fft == 512 ? launch_stft_nf<512>(p, wav, wav_stride, batch, X, mag, phase, s)
                             : launch_stft_nf<1024>(p, wav, wav_stride, batch, X, mag, phase, s);
}

template <int NF>
static int launch_istft_nf(const adv_plan* p, const float2* X, int64_t sb, int64_t st, int64_t sf, int batch,
                           float* out, double* stats, cudaStream_t s) {
    const Tiling tl = choose_tiling(p, batch);
    const size_t smem = IstftSmem<NF>::bytes(tl.hops_per_tile * p->d.hop);
    int rc = set_smem(istft_kernel<NF>, smem);
    if (rc != ADV_OK) return rc;
    dim3 grid(tl.tiles, batch);
    istft_kernel<NF><<<grid, kThreads, smem, s>>>(p->d, tl, X, sb, st, sf, out, stats);
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

int launch_istft(const adv_plan* p, const float2* X, int64_t sb, int64_t st, int64_t sf, int batch, float* out,
                 double* stats, cudaStream_t s) {
    return p->d.n_fft == 512 ? launch_istft_nf<512>(p, X, sb, st, sf, batch, out, stats, s)
                             : launch_istft_nf<1024>(p, X, sb, st, sf, batch, out, stats, s);
}

template <int NF, int MODE, bool FROM_SPEC>
static int launch_explain_inst(const adv_plan* p, const Tiling& tl, const float* wav, int64_t wav_stride,
                               const float2* X, int64_t sb, int64_t st, int64_t sf, const float* mask, int Fm,
                               int Tm, int batch, float* rel, float* irr, double* stats, cudaStream_t s) {
    const size_t smem = ExplainSmem<NF>::bytes(p->d.hop, tl.hops_per_tile * p->d.hop, FROM_SPEC);
    int rc = set_smem(explain_kernel<NF, MODE, FROM_SPEC>, smem);
    if (rc != ADV_OK) return rc;
    dim3 grid(tl.tiles, batch);
    explain_kernel<NF, MODE, FROM_SPEC><<<grid, kThreads, smem, s>>>(p->d, tl, wav, wav_stride, X, sb, st, sf,
                                                                     mask, Fm, Tm, rel, irr, stats);
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

int launch_explain(const adv_plan* p, const float* wav, int64_t wav_stride, const float2* X, int64_t sb,
                   int64_t st, int64_t sf, const float* mask, int Fm, int Tm, int mode, int batch, float* rel,
                   float* irr, double* stats, cudaStream_t s) {
    const Tiling tl = choose_tiling(p, batch);
    const bool spec = (X != nullptr);
#define ADV_EXPLAIN(NF, MODE, SPEC) \
    return launch_explain_inst<NF, MODE, SPEC>(p, tl, wav, wav_stride, X, sb, st, sf, mask, Fm, Tm, batch, rel, irr, stats, s)
    if (p->d.n_fft == 512) {
        if (mode == ADV_MASK_LOG1P) { if (spec) ADV_EXPLAIN(512, ADV_MASK_LOG1P, true); else ADV_EXPLAIN(512, ADV_MASK_LOG1P, false); }
        else { if (spec) ADV_EXPLAIN(512, ADV_MASK_LINEAR, true); else ADV_EXPLAIN(512, ADV_MASK_LINEAR, false); }
    } else {
        if (mode == ADV_MASK_LOG1P) { if (spec) ADV_EXPLAIN(1024, ADV_MASK_LOG1P, true); else ADV_EXPLAIN(1024, ADV_MASK_LOG1P, false); }
        else { if (spec) ADV_EXPLAIN(1024, ADV_MASK_LINEAR, true); else ADV_EXPLAIN(1024, ADV_MASK_LINEAR, false); }
    }
#undef ADV_EXPLAIN
}

}  // namespace adv
