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
// cooperates on one complex FFT (fft_core.cuh) and owns TWO ADJACENT frames (2u, 2u+1) of the tile:
//   * forward: both real frames ride in one complex FFT (split afterwards, register-local);
//   * inverse (istft): both Hermitian spectra ride in one complex inverse FFT;
//   * inverse (explain): the masked-in and masked-out spectra of ONE frame ride in one inverse FFT.
// The waveform segment of a tile is staged by one bulk-async copy (cp.async.bulk, the 1-D TMA path)
// completing on an mbarrier; tiles that touch the reflect-padded clip edges fall back to plain loads.
// Overlap-add is output-stationary and barrier-free: every unit accumulates its two frames into a
// PRIVATE shared-memory strip (hop + window-support samples); after one CTA barrier each output
// sample gathers the <= ceil(strip / 2hop)+1 strips that cover it, in a fixed order, multiplies by the
// precomputed reciprocal window envelope and is stored coalesced.  No atomics anywhere.
#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include "adv_internal.cuh"
#include "fft_core.cuh"
#include "transform_common.cuh"

namespace adv {

template <int NF>
struct Cfg {
    using G = Geo<NF>;
    static constexpr int UNITS = kThreads / G::LANES;
    static constexpr int FT = 2 * UNITS;  // frames per CTA
    static constexpr int MP = FT + 1;     // mask tile pitch (odd)
    // private overlap-add strip length (elements): hop + support, padded so that the two units of a
    // warp (n_fft 512) land in different banks
    static __host__ __device__ int strip(int hop, int support) {
        const int lb = hop + support;
        return G::LANES == 16 ? (((lb + 31) & ~31) + 16) : ((lb + 3) & ~3);
    }
    static __host__ __device__ size_t seg_floats(int hop) { return (size_t)(FT - 1) * hop + NF + 8; }
    static __host__ __device__ size_t strips_bytes(int hop, int support, bool from_spec) {
        size_t b = sizeof(float2) * UNITS * strip(hop, support);
        if (!from_spec && b < sizeof(float) * seg_floats(hop)) b = sizeof(float) * seg_floats(hop);
        return b;
    }
    static size_t stft_bytes(int hop) {
        return al16(sizeof(float2) * 32 * G::LANES) + al16(sizeof(float) * UNITS * G::SCRATCH) +
               al16(sizeof(float) * NF) + al16(sizeof(float) * seg_floats(hop)) + 16;
    }
    static size_t istft_bytes(int hop, int support, int tile_samples) {
        return al16(sizeof(float2) * 32 * G::LANES) + al16(sizeof(float) * UNITS * G::SCRATCH) +
               al16(sizeof(float) * NF) + al16(sizeof(float) * UNITS * strip(hop, support)) +
               al16(sizeof(double) * 2 * (kThreads / 32)) + al16(sizeof(float) * tile_samples);
    }
    static size_t explain_bytes(int hop, int support, bool from_spec, int tile_samples) {
        return al16(sizeof(float2) * 32 * G::LANES) + al16(sizeof(float) * UNITS * G::SCRATCH) +
               al16(sizeof(float) * NF) + al16(strips_bytes(hop, support, from_spec)) +
               al16(sizeof(double) * 4 * (kThreads / 32)) + al16(sizeof(float) * G::NBINS * MP) + 16 +
               al16(sizeof(float) * tile_samples);
    }
};

// ------------------------------------------------------------------------------------------------
// STFT
// ------------------------------------------------------------------------------------------------
#ifdef ADV_AB  // first-generation kernels: A/B builds only (-DADV_AB)
template <int NF, bool MAG, bool PHASE>
__global__ void __launch_bounds__(kThreads, 2)
stft_kernel(PlanDev P, const float* __restrict__ wav, int64_t wav_stride, float2* __restrict__ X,
            float* __restrict__ mag, float* __restrict__ phase) {
    using G = Geo<NF>;
    using C = Cfg<NF>;
    constexpr int UNITS = C::UNITS, FT = C::FT, F = G::NBINS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Carver cv{smem_raw};
    float2* tw_s = cv.take<float2>(32 * G::LANES);
    float* scratch = cv.take<float>(UNITS * G::SCRATCH);
    float* win_s = cv.take<float>(NF);
    float* seg = cv.take<float>(C::seg_floats(P.hop));
    uint64_t* bar = cv.take<uint64_t>(1);

    const int tid = threadIdx.x, b = blockIdx.y, t0 = blockIdx.x * FT;
    bool bulk;
    const int shift = stage_segment(seg, (FT - 1) * P.hop + NF, wav + (size_t)b * wav_stride, t0 * P.hop - NF / 2,
                                    P.n_in, bar, bulk);
    stage_tables<kThreads>(tw_s, 32 * G::LANES, win_s, NF, P);
    cp_async_wait_all();
    __syncthreads();
    stage_wait(bar, bulk);

    const int u = tid / G::LANES, l = tid % G::LANES;
    const int fa = t0 + 2 * u, fb = fa + 1;
    const float* sa = seg + shift + (2 * u) * P.hop + l;
    const float* sb = sa + P.hop;
    const float* wl = win_s + l;
    float2 v[32];
#pragma unroll
    for (int n1 = 0; n1 < 32; ++n1) {
        const float w = wl[n1 * G::R2];
        v[n1] = make_float2(sa[n1 * G::R2] * w, sb[n1 * G::R2] * w);
    }
    unit_fft_forward<NF>(v, l, TwSmem<G::LANES>{tw_s, l}, scratch + u * G::SCRATCH);
    float2 xa[17], xb[17];
    split_regs<NF>(v, l, xa, xb);

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
// STFT, persistent + software-pipelined (the production forward kernel)
//
// grid = min(tiles, resident CTAs); every CTA walks the tile list with stride gridDim.x.  The waveform
// segment of tile i+1 is requested (one bulk-async copy completing on an mbarrier) as soon as every
// thread holds tile i's samples in registers, so it lands while tile i's FFTs run: after the first tile
// no global-load latency is exposed, and the plan tables are staged once per CTA instead of once per
// tile.  Clip edges (torch.stft's centre=True reflect padding) and the <= 3 samples either side of the
// 16-byte aligned bulk range are filled with plain loads into the same buffer.
// ------------------------------------------------------------------------------------------------

template <int NF>
struct PCfg {  // shared-memory carve-up of the persistent kernels
    using G = Geo<NF>;
    static constexpr int UNITS = kThreads / G::LANES, FT = 2 * UNITS;
    static __host__ __device__ size_t seg_floats(int hop) { return (size_t)(FT - 1) * hop + NF + 8; }
    static size_t stft_bytes(int hop) {
        return al16(sizeof(float2) * 32 * G::LANES) + al16(sizeof(float) * UNITS * GeoV<NF>::SCRATCH) +
               al16(sizeof(float) * NF) + al16(sizeof(float) * seg_floats(hop)) + 16;
    }
};

template <int NF, bool MAG, bool PHASE, bool RECT>
__global__ void __launch_bounds__(kThreads, 2)
stft_p_kernel(PlanDev P, const float* __restrict__ wav, int64_t wav_stride, int total_tiles, int tiles_per_clip,
              float2* __restrict__ X, float* __restrict__ mag, float* __restrict__ phase) {
    using G = Geo<NF>;
    using C = PCfg<NF>;
    constexpr int UNITS = C::UNITS, FT = C::FT, F = G::NBINS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Carver cv{smem_raw};
    float2* tw_s = cv.take<float2>(32 * G::LANES);
    float* scratch = cv.take<float>(UNITS * GeoV<NF>::SCRATCH);
    float* win_s = cv.take<float>(NF);
    float* seg = cv.take<float>(C::seg_floats(P.hop));
    uint64_t* bar = cv.take<uint64_t>(1);

    const int tid = threadIdx.x, u = tid / G::LANES, l = tid % G::LANES;
    const int seglen = (FT - 1) * P.hop + NF;
    int tile = blockIdx.x;
    if (tid == 0) mbar_init(bar, 1);
    stage_tables<kThreads>(tw_s, 32 * G::LANES, win_s, NF, P);
    __syncthreads();  // barrier initialised before anyone arms or polls it
    int b = tile / tiles_per_clip, t0 = (tile - b * tiles_per_clip) * FT;
    int shift = stage_segment_async<kThreads>(seg, seglen, wav + (size_t)b * wav_stride, t0 * P.hop - NF / 2, P.n_in, bar);
    cp_async_wait_all();

    float* my = scratch + u * GeoV<NF>::SCRATCH;
    const TwSmem<G::LANES> tw{tw_s, l};
    for (uint32_t it = 0;; ++it) {
        __syncthreads();  // plain-load part of the segment (and, first time round, the tables) visible
        mbar_wait(bar, it & 1);
        float2 v[32];
        {
            const float* sa = seg + shift + (2 * u) * P.hop + l;
            const float* sb = sa + P.hop;
            const float* wl = win_s + l;
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) {
                if (RECT) {
                    v[n1] = make_float2(sa[n1 * G::R2], sb[n1 * G::R2]);
                } else {
                    const float w = wl[n1 * G::R2];
                    v[n1] = make_float2(sa[n1 * G::R2] * w, sb[n1 * G::R2] * w);
                }
            }
        }
        const int cur_b = b, fa = t0 + 2 * u;
        const int next = tile + gridDim.x;
        __threadfence_block();  // the sample loads have returned: a sync alone does not wait for queued shared-memory loads
        __syncthreads();  // every thread holds its samples: the buffer is free for the next tile
        if (next < total_tiles) {
            b = next / tiles_per_clip;
            t0 = (next - b * tiles_per_clip) * FT;
            shift = stage_segment_async<kThreads>(seg, seglen, wav + (size_t)b * wav_stride, t0 * P.hop - NF / 2,
                                                  P.n_in, bar);
        }
        unit_fft_forward_v<NF>(v, l, tw, my);
        float2 xa[17], xb[17];
        split_regs<NF>(v, l, xa, xb);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int t = fa + half;
            if (t >= P.T) continue;
            const float2* x = half ? xb : xa;
            const size_t row = ((size_t)cur_b * P.T + t) * F;
#pragma unroll
            for (int i = 0; i < 17; ++i) {
                const int bin = bin_of<NF>(l, i);
                if (bin < 0) continue;
                X[row + bin] = x[i];
                if (MAG) {
                    const float r2 = fmaf(x[i].x, x[i].x, x[i].y * x[i].y);
                    mag[row + bin] = sqrtf(r2);
                }
                if (PHASE) phase[row + bin] = fast_atan2f(x[i].y, x[i].x);
            }
        }
        if (next >= total_tiles) break;
        tile = next;
    }
}

#endif  // ADV_AB
// ------------------------------------------------------------------------------------------------
// STFT, warp-autonomous: every warp owns its frames (4 for n_fft 512, 2 for 1024), its slice of the
// waveform, its mbarrier and its place in the item list - no CTA barrier in steady state, so the warps of
// an SM drift into different phases (shared-memory loads / FMA butterflies / global stores) and the three
// pipes overlap instead of being hammered in lock-step.
// n_fft 512: the two units of a warp read slices whose starts differ by 2*hop samples; when that is a
// multiple of 32 floats their loads would collide bank for bank, so the odd unit reads its samples rotated
// by one radix-32 row (sample n1+1 where the even unit reads n1) and uses a twiddle table that carries the
// compensating phase exp(-2 pi i k1 / 32) (second half of PlanDev::tw) - conflict-free at no cost.
// ------------------------------------------------------------------------------------------------
template <int NF>
struct WCfg {
    using G = Geo<NF>;
    static constexpr int WARPS = kThreads / 32, UPW = 32 / G::LANES, FW = 2 * UPW;  // units, frames per warp
    static __host__ __device__ size_t seg_floats(int hop) { return ((size_t)(FW - 1) * hop + NF + 8 + 3) & ~size_t(3); }
    // per-unit scratch in floats: the transposes need GeoV::SCRATCH; with BULK stores the unit's two finished
    // rows (2 * NBINS complex + 16 bytes of alignment slack) are staged in the same memory.  Stride = 16 mod 32
    // floats so that the two units of a warp sit in opposite bank halves.
    static constexpr int unit_floats(bool bulk) {
        int n = GeoV<NF>::SCRATCH;
        const int need = 4 * G::NBINS + 4;
        if (bulk && need > n) n = need;
        return ((n + 15) / 32) * 32 + 16;
    }
    static size_t stft_bytes(int hop, bool bulk) {
        return al16(sizeof(float2) * 2 * 32 * G::LANES) + al16(sizeof(float) * WARPS * UPW * unit_floats(bulk)) +
               al16(sizeof(float) * NF) + al16(sizeof(float) * WARPS * seg_floats(hop)) + al16(8 * WARPS);
    }
};

template <int NF, bool MAG, bool PHASE, bool RECT, bool BULK, bool ZP = false>  // ZP: zero instead of reflect padding
__global__ void __launch_bounds__(kThreads, 2)
stft_w_kernel(PlanDev P, const float* __restrict__ wav, int64_t wav_stride, int total_items, int items_per_clip,
              float2* __restrict__ X, float* __restrict__ mag, float* __restrict__ phase) {
    using G = Geo<NF>;
    using C = WCfg<NF>;
    constexpr int F = G::NBINS, FW = C::FW;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Carver cv{smem_raw};
    float2* tw_s = cv.take<float2>(2 * 32 * G::LANES);
    constexpr int UF = C::unit_floats(BULK);
    float* scratch = cv.take<float>(C::WARPS * C::UPW * UF);
    float* win_s = cv.take<float>(NF);
    float* seg_all = cv.take<float>(C::WARPS * C::seg_floats(P.hop));
    uint64_t* bars = cv.take<uint64_t>(C::WARPS);

    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int uw = lane / G::LANES, l = lane % G::LANES;  // unit inside the warp, lane inside the unit
    float* seg = seg_all + (size_t)w * C::seg_floats(P.hop);
    uint64_t* bar = bars + w;
    const int seglen = (FW - 1) * P.hop + NF;
    const int stride = gridDim.x * C::WARPS;
    int item = blockIdx.x * C::WARPS + w;

    if (lane == 0) mbar_init(bar, 1);
    for (int i = tid; i < 32 * G::LANES; i += kThreads) cp_async16(tw_s + 2 * i, P.tw + 2 * i);  // both tables
    for (int i = tid; i < NF / 4; i += kThreads) cp_async16(win_s + 4 * i, P.window + 4 * i);
    pdl_launch_dependents();
    cp_async_wait_all();
    __syncthreads();  // tables staged, barriers initialised
    pdl_wait();       // everything above touches plan constants only
    if (item >= total_items) return;

    // bank rotation of the odd unit (see above): only when the units' slices start on the same bank
    const int rot = (G::LANES == 16 && ((2 * P.hop) & 31) == 0) ? uw : 0;
    float* my = scratch + (w * C::UPW + uw) * UF;
    const bool elected = (l == 0);
    const TwSmem<G::LANES> tw{tw_s + rot * 32 * G::LANES, l};
    int b = item / items_per_clip, t0 = (item - b * items_per_clip) * FW;
    int shift = stage_segment_async<32>(seg, seglen, wav + (size_t)b * wav_stride, t0 * P.hop - NF / 2, P.n_in, bar, lane, ZP);

    for (uint32_t it = 0;; ++it) {
        __syncwarp();  // plain-load part of the slice visible to the warp
        mbar_wait(bar, it & 1);
        float2 v[32];
        {
            const float* sa = seg + shift + (2 * uw) * P.hop + l + rot * G::R2;
            const float* sb = sa + P.hop;
            const float* wl = win_s + l + rot * G::R2;
            const int wrap = rot * NF;  // the rotated unit's last row is row 0
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) {
                const int o = n1 * G::R2 - (n1 == 31 ? wrap : 0);
                if (RECT) {
                    v[n1] = make_float2(sa[o], sb[o]);
                } else {
                    const float ww = wl[o];
                    v[n1] = make_float2(sa[o] * ww, sb[o] * ww);
                }
            }
        }
        const int cur_b = b, fa = t0 + 2 * uw;
        const int next = item + stride;
        __threadfence_block();  // the sample loads have returned: a sync alone does not wait for queued shared-memory loads
        __syncwarp();  // every lane holds its samples: the slice buffer is free for the next item
        if (next < total_items) {
            b = next / items_per_clip;
            t0 = (next - b * items_per_clip) * FW;
            shift = stage_segment_async<32>(seg, seglen, wav + (size_t)b * wav_stride, t0 * P.hop - NF / 2, P.n_in, bar,
                                            lane, ZP);
        }
        if (BULK) {  // the previous item's rows have left the staging memory (it doubles as transpose scratch)
            if (elected) bulk_wait_read();
        }
        unit_fft_forward_v<NF>(v, l, tw, my);
        float2 xa[17], xb[17];
        split_regs<NF>(v, l, xa, xb);
        if (BULK) {
            // rows fa, fa+1 of one clip are adjacent in X / mag / phase: stage both and hand each array to one
            // bulk-async store (full-line writes, no per-lane store traffic through the LSU)
            const int nrows = min(2, P.T - fa);
            if (nrows > 0) {
                const size_t row = ((size_t)cur_b * P.T + fa) * F;
                unsigned char* stage = reinterpret_cast<unsigned char*>(my);
                {
                    unsigned char* g = reinterpret_cast<unsigned char*>(X + row);
                    float2* st = reinterpret_cast<float2*>(stage + (reinterpret_cast<uintptr_t>(g) & 15));
                    __syncwarp();  // transposes of this unit done with the scratch
#pragma unroll
                    for (int i = 0; i < 17; ++i) {
                        const int bin = bin_of<NF>(l, i);
                        if (bin < 0) continue;
                        st[bin] = xa[i];
                        st[F + bin] = xb[i];
                    }
                    fence_async_smem();
                    __syncwarp();
                    unit_store_bulk(stage, g, (uint32_t)nrows * F * 8u, l, elected);
                }
                if (MAG || PHASE) {
#pragma unroll
                    for (int i = 0; i < 17; ++i) {  // (re, im) -> (|X|, angle X) in place
                        const float2 a = xa[i], c = xb[i];
                        xa[i] = make_float2(sqrtf(fmaf(a.x, a.x, a.y * a.y)), PHASE ? fast_atan2f(a.y, a.x) : 0.0f);
                        xb[i] = make_float2(sqrtf(fmaf(c.x, c.x, c.y * c.y)), PHASE ? fast_atan2f(c.y, c.x) : 0.0f);
                    }
                    if (elected) bulk_wait_read();
                    __syncwarp();
                    // first half of the staging area: magnitudes, second half: phases (each 2 rows of F floats)
                    unsigned char* gm = reinterpret_cast<unsigned char*>(mag + row);
                    unsigned char* gp = reinterpret_cast<unsigned char*>(phase + row);
                    unsigned char* stage_p = stage + ((2 * F * 4 + 16 + 15) & ~15);
                    float* sm_ = reinterpret_cast<float*>(stage + (MAG ? (reinterpret_cast<uintptr_t>(gm) & 15) : 0));
                    float* sp_ = reinterpret_cast<float*>(stage_p + (PHASE ? (reinterpret_cast<uintptr_t>(gp) & 15) : 0));
#pragma unroll
                    for (int i = 0; i < 17; ++i) {
                        const int bin = bin_of<NF>(l, i);
                        if (bin < 0) continue;
                        if (MAG) { sm_[bin] = xa[i].x; sm_[F + bin] = xb[i].x; }
                        if (PHASE) { sp_[bin] = xa[i].y; sp_[F + bin] = xb[i].y; }
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (MAG) unit_store_bulk(stage, gm, (uint32_t)nrows * F * 4u, l, elected);
                    if (PHASE) unit_store_bulk(stage_p, gp, (uint32_t)nrows * F * 4u, l, elected);
                }
            }
        } else {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int t = fa + half;
            if (t >= P.T) continue;
            const float2* x = half ? xb : xa;
            const size_t row = ((size_t)cur_b * P.T + t) * F;
#pragma unroll
            for (int i = 0; i < 17; ++i) {
                const int bin = bin_of<NF>(l, i);
                if (bin < 0) continue;
                X[row + bin] = x[i];
                if (MAG) mag[row + bin] = fast_abs2(x[i]);
                if (PHASE) phase[row + bin] = fast_atan2f(x[i].y, x[i].x);
            }
        }
        }
        if (next >= total_items) break;
        item = next;
    }
    if (BULK) {  // shared memory must outlive the reads of the last bulk stores
        if (elected) bulk_wait_read();
    }
}

// ------------------------------------------------------------------------------------------------
// shared pieces of the overlap-add kernels
// ------------------------------------------------------------------------------------------------

// Sum of the private strips that cover offset x (samples past the first strip's origin).  Strip u
// starts at 2*hop*u and holds `lb` elements.  Fixed (descending-u) order => bit-reproducible.
template <int UNITS>
__device__ __forceinline__ float gather1(const float* pb, int strip, int lb, int two_hop, float inv_two_hop, int x) {
    int u = min(UNITS - 1, (int)(((float)x + 0.5f) * inv_two_hop));
    int k = x - u * two_hop;
    float acc = 0.0f;
    while (u >= 0 && k < lb) {
        acc += pb[u * strip + k];
        --u;
        k += two_hop;
    }
    return acc;
}
template <int UNITS>
__device__ __forceinline__ float2 gather2(const float2* pb, int strip, int lb, int two_hop, float inv_two_hop, int x) {
    int u = min(UNITS - 1, (int)(((float)x + 0.5f) * inv_two_hop));
    int k = x - u * two_hop;
    float2 acc = make_float2(0.f, 0.f);
    while (u >= 0 && k < lb) {
        const float2 c = pb[u * strip + k];
        acc.x += c.x;
        acc.y += c.y;
        --u;
        k += two_hop;
    }
    return acc;
}

// ------------------------------------------------------------------------------------------------
// iSTFT
// ------------------------------------------------------------------------------------------------
template <int NF>
__global__ void __launch_bounds__(kThreads, 2)
istft_kernel(PlanDev P, Tiling TL, const float2* __restrict__ X, int64_t sb, int64_t st, int64_t sf,
             float* __restrict__ out, double* __restrict__ stats) {
    using G = Geo<NF>;
    using C = Cfg<NF>;
    constexpr int UNITS = C::UNITS;
    const int support = P.whi - P.wlo, lb = P.hop + support, strip = C::strip(P.hop, support);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Carver cv{smem_raw};
    float2* tw_s = cv.take<float2>(32 * G::LANES);
    float* scratch = cv.take<float>(UNITS * G::SCRATCH);
    float* win_s = cv.take<float>(NF);
    float* pb = cv.take<float>(UNITS * strip);
    double* red = cv.take<double>(2 * (kThreads / 32));
    float* env_s = cv.take<float>(TL.hops_per_tile * P.hop);

    const int tid = threadIdx.x, b = blockIdx.y;
    const TileGeom g = tile_geom<NF>(P, TL, blockIdx.x);
    const int S = g.s1 - g.s0;
    stage_tables<kThreads>(tw_s, 32 * G::LANES, win_s, NF, P);
    const int u = tid / G::LANES, l = tid % G::LANES;
    const int fa = g.t_lo + 2 * u, fb = fa + 1;
    const bool va = fa <= g.t_hi, vb = fb <= g.t_hi;

    float2 ya[17], yb[17];
    {
        const float2* xa_p = X + (size_t)b * sb + (size_t)fa * st;
        const float2* xb_p = xa_p + st;
#pragma unroll
        for (int i = 0; i < 17; ++i) {
            const int bin = bin_of<NF>(l, i);
            const float2 z = make_float2(0.f, 0.f);
            ya[i] = (va && bin >= 0) ? __ldg(xa_p + (size_t)bin * sf) : z;
            yb[i] = (vb && bin >= 0) ? __ldg(xb_p + (size_t)bin * sf) : z;
        }
    }
    cp_async_wait_all();
    __syncthreads();
    stage_env<kThreads>(env_s, P.inv_env + g.s0, S);  // needed only by the epilogue

    float2 v[32];
    merge_regs<NF>(v, l, ya, yb);
    unit_fft_inverse<NF>(v, l, TwSmem<G::LANES>{tw_s, l}, scratch + u * G::SCRATCH);

    // private strip of the unit: frame a at [0, support), frame b at [hop, hop + support)
    {
        float* pbu = pb + u * strip;
        const float* wl = win_s + l;
        const int c0 = l - P.wlo;
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
            const int k = n1 * G::R2 + c0;
            if ((unsigned)k < (unsigned)support) pbu[k] = v[n1].x * wl[n1 * G::R2];
        }
        const int ovl = support - P.hop;  // samples of frame b that land on frame a's
        if (ovl < 0)  // degenerate (hop > support, only legal for single-frame plans): clear the gap
            for (int k = support + l; k < P.hop; k += G::LANES) pbu[k] = 0.0f;
        __syncwarp();
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
            const int k = n1 * G::R2 + c0;
            if ((unsigned)k < (unsigned)support) {
                const float old = k < ovl ? pbu[P.hop + k] : 0.0f;
                pbu[P.hop + k] = fmaf(v[n1].y, wl[n1 * G::R2], old);
            }
        }
    }
    cp_async_wait_all();
    __syncthreads();

    double acc[2] = {0.0, 0.0};
    float* orow = out + (size_t)b * P.n_out + g.s0;
    const int two_hop = 2 * P.hop, x0 = g.p0 - (g.t_lo * P.hop + P.wlo);
    const float inv_two_hop = 1.0f / (float)two_hop;
    for (int q = tid; q < S; q += kThreads) {
        const float y = gather1<UNITS>(pb, strip, lb, two_hop, inv_two_hop, x0 + q) * env_s[q];
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
// iSTFT, persistent + software-pipelined (the production inverse kernel)
//
// grid = min(tiles, resident CTAs); every CTA walks the tile list.  Differences to istft_kernel:
//   * the next tile's spectrum rows are requested (plain loads into registers that are dead between the
//     strip write and the next merge) BEFORE the barrier and the gather epilogue of the current tile, so
//     their DRAM latency is covered by the epilogue instead of being exposed at tile start;
//   * the gather epilogue handles VEC consecutive samples per step (128-/64-bit strip loads and output
//     stores; strip index arithmetic once per group) - it was a third of the old kernel's instructions;
//   * the strip origin is the window-support start rounded DOWN to a multiple of VEC (the extra leading
//     taps are exact zeros of the window table), which keeps every strip access VEC-aligned;
//   * plan tables are staged once per CTA; the reciprocal envelope is read straight from L2.
// CONTIG: spectrum rows are contiguous (bin stride 1): offsets become immediates.
// ------------------------------------------------------------------------------------------------

template <int NF>
struct IPCfg {
    using G = Geo<NF>;
    static constexpr int UNITS = kThreads / G::LANES;
    // strip pitch: hop + support (support measured from the VEC-aligned origin, <= 3 more taps), padded like Cfg
    static __host__ __device__ int strip(int hop, int support) { return Cfg<NF>::strip(hop, support + 4); }
    static size_t bytes(int hop, int support, int tile_samples) {
        return al16(sizeof(float2) * 32 * G::LANES) + al16(sizeof(float) * UNITS * GeoV<NF>::SCRATCH) +
               al16(sizeof(float) * NF) + al16(sizeof(float) * UNITS * strip(hop, support)) +
               al16(sizeof(double) * 2 * (kThreads / 32)) + al16(sizeof(float) * tile_samples);
    }
};

template <int NF, int VEC, bool CONTIG>
__global__ void __launch_bounds__(kThreads, 2)
istft_p_kernel(PlanDev P, Tiling TL, int total_tiles, const float2* __restrict__ X, int64_t sb, int64_t st, int64_t sf,
               float* __restrict__ out, double* __restrict__ stats) {
    using G = Geo<NF>;
    using C = IPCfg<NF>;
    using V = typename VecT<VEC>::T;
    constexpr int UNITS = C::UNITS;
    const int wlo = P.wlo & ~(VEC - 1);
    const int support = P.whi - wlo, lb = P.hop + support, strip = C::strip(P.hop, P.whi - P.wlo);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Carver cv{smem_raw};
    float2* tw_s = cv.take<float2>(32 * G::LANES);
    float* scratch = cv.take<float>(UNITS * GeoV<NF>::SCRATCH);
    float* win_s = cv.take<float>(NF);
    float* pb = cv.take<float>(UNITS * strip);
    double* red = cv.take<double>(2 * (kThreads / 32));
    float* env_s = cv.take<float>(TL.hops_per_tile * P.hop);

    const int tid = threadIdx.x, u = tid / G::LANES, l = tid % G::LANES;
    stage_tables<kThreads>(tw_s, 32 * G::LANES, win_s, NF, P);
    int tile = blockIdx.x;
    int b = tile / TL.tiles;
    TileGeom g = tile_geom<NF>(P, TL, tile - b * TL.tiles);
    float* my = scratch + u * GeoV<NF>::SCRATCH;
    // lane-private slots inside the unit's transpose scratch (dead between the last transpose read and the
    // next tile's first transpose write): frame a's bins arrive there by cp.async, frame b's in registers
    float2* slot = reinterpret_cast<float2*>(my) + l * 17;
    static_assert(17 * 8 * G::LANES <= GeoV<NF>::SCRATCH * 4, "row staging must fit the transpose scratch");

    float2 yb[17];
    bool va_cur = false;
    auto load_rows = [&](int bb, const TileGeom& gg) -> bool {
        const int fa = gg.t_lo + 2 * u;
        const bool va = fa <= gg.t_hi, vb = fa + 1 <= gg.t_hi;
        const float2* xa_p = X + (size_t)bb * sb + (size_t)fa * st;
        const float2* xb_p = xa_p + st;
#pragma unroll
        for (int i = 0; i < 17; ++i) {
            const int bin = bin_of<NF>(l, i);
            const float2 z = make_float2(0.f, 0.f);
            const size_t o = CONTIG ? (size_t)bin : (size_t)bin * sf;
            if (va && bin >= 0)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(slot + i)), "l"(xa_p + o) : "memory");
            yb[i] = (vb && bin >= 0) ? __ldg(xb_p + o) : z;
        }
        return va;
    };
    va_cur = load_rows(b, g);
    stage_env<kThreads>(env_s, P.inv_env + g.s0, g.s1 - g.s0);
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();

    float* pbu = pb + u * strip;
    const TwSmem<G::LANES> tw{tw_s, l};
    const float* wl = win_s + l;
    const int c0 = l - wlo, ovl = support - P.hop, two_hop = 2 * P.hop;
    for (;;) {
        const int next = tile + gridDim.x;
        const int cur_b = b, cur_tile = tile - b * TL.tiles;
        const TileGeom cg = g;
        // warps whose units are all past the tile's last frame have nothing to transform (short tiles are chosen
        // on purpose when they balance the rounds better, see choose_tiling)
        const bool warp_live = __ballot_sync(0xffffffffu, va_cur) != 0;
        if (warp_live) {
        float2 v[32];
        {
            float2 ya[17];
#pragma unroll
            for (int i = 0; i < 17; ++i)
                ya[i] = (va_cur && bin_of<NF>(l, i) >= 0) ? slot[i] : make_float2(0.f, 0.f);
            merge_regs<NF>(v, l, ya, yb);
        }
        unit_fft_inverse_v<NF>(v, l, tw, my);  // (starts with a __syncwarp: the unit's slot reads are done)
        // private strip of the unit: frame a at [0, support), frame b at [hop, hop + support)
        // (the previous tile's epilogue finished reading the strips: barrier at the loop end)
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
            const int k = n1 * G::R2 + c0;
            if ((unsigned)k < (unsigned)support) pbu[k] = v[n1].x * wl[n1 * G::R2];
        }
        if (ovl < 0)  // degenerate (hop > support, only legal for single-frame plans): clear the gap
            for (int k = support + l; k < P.hop; k += G::LANES) pbu[k] = 0.0f;
        __syncwarp();
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
            const int k = n1 * G::R2 + c0;
            if ((unsigned)k < (unsigned)support) {
                const float old = k < ovl ? pbu[P.hop + k] : 0.0f;
                pbu[P.hop + k] = fmaf(v[n1].y, wl[n1 * G::R2], old);
            }
        }
        } else {
            // the gather may still touch this strip (frames clipped at the end of the clip): it must read zeros
            for (int k = l; k < lb; k += G::LANES) pbu[k] = 0.0f;
        }
        if (next < total_tiles) {  // next tile's rows: in flight across the barrier and the epilogue
            b = next / TL.tiles;
            g = tile_geom<NF>(P, TL, next - b * TL.tiles);
            __syncwarp();  // every lane of the unit is past its last transpose read
            va_cur = load_rows(b, g);
            cp_async_commit();
            cp_async_wait_group<1>();  // this tile's envelope (older group) has landed; the rows may still fly
        } else {
            cp_async_wait_all();
        }
        __syncthreads();

        // gather: VEC samples per step; strips covering offset x are u = x / two_hop, u-1, ... while k < lb
        const int S = cg.s1 - cg.s0;
        const int x0 = cg.p0 - (cg.t_lo * P.hop + wlo);
        float* orow = out + (size_t)cur_b * P.n_out + cg.s0;
        double acc0 = 0.0, acc1 = 0.0;
        const float inv_two_hop = 1.0f / (float)two_hop;
        for (int q = tid * VEC; q < S; q += kThreads * VEC) {
            const int x = x0 + q;
            int uu = min(UNITS - 1, (int)(((float)x + 0.5f) * inv_two_hop));  // exact: x < 2^20
            int k = x - uu * two_hop;
            V a = vzero<V>();
            while (uu >= 0 && k < lb) {
                vadd(a, *reinterpret_cast<const V*>(pb + uu * strip + k));
                --uu;
                k += two_hop;
            }
            if (VEC == 1 || q + VEC <= S) {
                const V e = *reinterpret_cast<const V*>(env_s + q);
                float sq;
                const float sm = vmul_stats(a, e, sq);
                *reinterpret_cast<V*>(orow + q) = a;
                acc0 += (double)sm;
                acc1 += (double)sq;
            } else {  // ragged end of the clip (n_out not a multiple of VEC): element by element
                const float* af = reinterpret_cast<const float*>(&a);
                for (int j = 0; j < VEC && q + j < S; ++j) {
                    const float y = af[j] * env_s[q + j];
                    orow[q + j] = y;
                    acc0 += (double)y;
                    acc1 += (double)y * (double)y;
                }
            }
        }
        if (stats != nullptr) {
            double acc[2] = {acc0, acc1};
            block_sum<2>(acc, red);
            if (tid == 0) {
                double* srow = stats + ((size_t)cur_b * TL.tiles + cur_tile) * 2;
                srow[0] = acc[0];
                srow[1] = acc[1];
            }
        }
        if (next >= total_tiles) break;
        tile = next;
        __syncthreads();  // strips, envelope tile and reduction scratch are free again
        stage_env<kThreads>(env_s, P.inv_env + g.s0, g.s1 - g.s0);  // lands during the FFTs
        cp_async_commit();
        cp_async_wait_group<1>();  // own row copies (issued before the epilogue, long since landed)
    }
}

// ------------------------------------------------------------------------------------------------
// fused explain: wave (or spectrum) + mask -> masked-in / masked-out waveforms
// ------------------------------------------------------------------------------------------------

template <int NF, int MODE, bool FROM_SPEC>
__global__ void __launch_bounds__(kThreads, 1)
explain_kernel(PlanDev P, Tiling TL, const float* __restrict__ wav, int64_t wav_stride,
               const float2* __restrict__ X, int64_t sb, int64_t st, int64_t sf,
               const float* __restrict__ mask, int Fm, int Tm, int drop,
               float* __restrict__ rel, float* __restrict__ irr, double* __restrict__ stats) {
    using G = Geo<NF>;
    using C = Cfg<NF>;
    constexpr int UNITS = C::UNITS, FT = C::FT, MP = C::MP, F = G::NBINS;
    const int support = P.whi - P.wlo, lb = P.hop + support, strip = C::strip(P.hop, support);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Carver cv{smem_raw};
    float2* tw_s = cv.take<float2>(32 * G::LANES);
    float* scratch = cv.take<float>(UNITS * G::SCRATCH);
    float* win_s = cv.take<float>(NF);
    // .x masked-in, .y masked-out; the staged waveform shares this memory (dead once in registers)
    float2* pb = reinterpret_cast<float2*>(cv.take<unsigned char>(C::strips_bytes(P.hop, support, FROM_SPEC)));
    float* seg = reinterpret_cast<float*>(pb);
    double* red = cv.take<double>(4 * (kThreads / 32));
    float* mask_s = cv.take<float>(F * MP);
    uint64_t* bar = cv.take<uint64_t>(1);
    float* env_s = cv.take<float>(TL.hops_per_tile * P.hop);

    const int tid = threadIdx.x, b = blockIdx.y;
    const TileGeom g = tile_geom<NF>(P, TL, blockIdx.x);
    const int S = g.s1 - g.s0;
    stage_env<kThreads>(env_s, P.inv_env + g.s0, S);  // needed only by the epilogue
    bool bulk = false;
    int shift = 0;
    if (!FROM_SPEC)
        shift = stage_segment(seg, (FT - 1) * P.hop + NF, wav + (size_t)b * wav_stride, g.t_lo * P.hop - NF / 2,
                              P.n_in, bar, bulk);
    {   // mask tile [F][FT] <- mask[b][f][t_lo + c], zero outside the mask / tile
        const float* mrow = mask + (size_t)b * Fm * Tm;
#pragma unroll 8
        for (int e = tid; e < F * FT; e += kThreads) {
            const int f = e / FT, c = e % FT;
            const int t = g.t_lo + c;
            mask_s[f * MP + c] = (f < Fm && t < Tm && t <= g.t_hi) ? __ldg(mrow + (size_t)f * Tm + t) : 0.0f;
        }
    }
    for (int i = tid; i < 32 * G::LANES; i += kThreads) tw_s[i] = P.tw[i];
    for (int i = tid; i < NF; i += kThreads) win_s[i] = P.window[i];
    __syncthreads();

    const int u = tid / G::LANES, l = tid % G::LANES;
    float* my = scratch + u * G::SCRATCH;
    const int fa = g.t_lo + 2 * u;
    const float* wl = win_s + l;

    float2 v[32];
    float2 xa[17], xb[17];
    if (FROM_SPEC) {
        const bool va = fa <= g.t_hi, vb = fa + 1 <= g.t_hi;
        const float2* xa_p = X + (size_t)b * sb + (size_t)fa * st;
        const float2* xb_p = xa_p + st;
#pragma unroll
        for (int i = 0; i < 17; ++i) {
            const int bin = bin_of<NF>(l, i);
            const float2 z = make_float2(0.f, 0.f);
            xa[i] = (va && bin >= 0) ? __ldg(xa_p + (size_t)bin * sf) : z;
            xb[i] = (vb && bin >= 0) ? __ldg(xb_p + (size_t)bin * sf) : z;
        }
    } else {
        stage_wait(bar, bulk);
        const float* sa = seg + shift + (2 * u) * P.hop + l;
        const float* sbp = sa + P.hop;
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
            const float w = wl[n1 * G::R2];
            v[n1] = make_float2(sa[n1 * G::R2] * w, sbp[n1 * G::R2] * w);
        }
        __syncthreads();  // every thread holds its samples: the segment's memory becomes the strips
        unit_fft_forward<NF>(v, l, TwSmem<G::LANES>{tw_s, l}, my);
        split_regs<NF>(v, l, xa, xb);
    }

    float2* pbu = pb + u * strip;
    const int c0 = l - P.wlo, ovl = support - P.hop;
    if (ovl < 0)  // degenerate (hop > support, only legal for single-frame plans): clear the gap
        for (int k = support + l; k < P.hop; k += G::LANES) pbu[k] = make_float2(0.f, 0.f);
    // the two frames of the unit go through the same code: frame b's spectrum is rotated into xa
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        const int t = fa + half;
        const bool valid = t <= g.t_hi;
        const int col = 2 * u + half;
        {
            float2 yr[17], yi[17];
#pragma unroll
            for (int i = 0; i < 17; ++i) {
                const int bin = bin_of<NF>(l, i);
                const bool live = bin >= 0 && valid;
                const float m = live ? mask_s[bin * MP + col] : 0.0f;
                float gr, gi;
                mask_gains<MODE>(xa[i], m, gr, gi);
                yr[i] = live ? make_float2(xa[i].x * gr, xa[i].y * gr) : make_float2(0.f, 0.f);
                // out-of-mask bins: zero extension sends them to the masked-out branch; `drop` removes them
                const bool keep_i = live && !(drop && (bin >= Fm || t >= Tm));
                yi[i] = keep_i ? make_float2(xa[i].x * gi, xa[i].y * gi) : make_float2(0.f, 0.f);
            }
            merge_regs<NF>(v, l, yr, yi);
        }
#pragma unroll
        for (int i = 0; i < 17; ++i) xa[i] = xb[i];
        // (starts with a __syncwarp: frame a's strip stores below are visible to the whole unit
        //  before frame b's read-modify-write)
        unit_fft_inverse<NF>(v, l, TwSmem<G::LANES>{tw_s, l}, my);
        // v[n1] = n_fft * (rel[n], irr[n]) of frame t -> private strip at [half*hop, half*hop + support)
        float2* dst = pbu + half * P.hop;
        const int keep = half ? ovl : 0;  // leading samples that add onto frame a's
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
            const int k = n1 * G::R2 + c0;
            if ((unsigned)k < (unsigned)support) {
                const float w = wl[n1 * G::R2];
                float2 o = k < keep ? dst[k] : make_float2(0.f, 0.f);
                o.x = fmaf(v[n1].x, w, o.x);
                o.y = fmaf(v[n1].y, w, o.y);
                dst[k] = o;
            }
        }
    }
    cp_async_wait_all();
    __syncthreads();

    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    float* rrow = rel + (size_t)b * P.n_out + g.s0;
    float* irow = irr + (size_t)b * P.n_out + g.s0;
    const int two_hop = 2 * P.hop, x0 = g.p0 - (g.t_lo * P.hop + P.wlo);
    const float inv_two_hop = 1.0f / (float)two_hop;
    for (int q = tid; q < S; q += kThreads) {
        const float e = env_s[q];
        const float2 o = gather2<UNITS>(pb, strip, lb, two_hop, inv_two_hop, x0 + q);
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
// fused explain, n_fft = 512, "wide" units: 512 threads = 16 warps, one warp per FFT (fft_core.cuh w512).
// Same tiling, staging, strips and gather as explain_kernel<512>; half the registers per thread, so
// twice the warps are resident per SM (the narrow-unit kernel is FP32-issue / latency bound at 8 warps).
// ------------------------------------------------------------------------------------------------
constexpr int kWideThreads = 512;

struct WideCfg {
    static constexpr int UNITS = 16, FT = 32, MP = FT + 1, F = 257, NF = 512;
    static __host__ __device__ int strip(int hop, int support) { return (hop + support + 3) & ~3; }
    static __host__ __device__ size_t seg_floats(int hop) { return (size_t)(FT - 1) * hop + NF + 8; }
    static __host__ __device__ size_t strips_bytes(int hop, int support, bool from_spec) {
        size_t b = sizeof(float2) * UNITS * strip(hop, support);
        if (!from_spec && b < sizeof(float) * seg_floats(hop)) b = sizeof(float) * seg_floats(hop);
        return b;
    }
    static size_t bytes(int hop, int support, bool from_spec, int tile_samples) {
        return al16(sizeof(float2) * 512) + al16(sizeof(float) * UNITS * w512::SCRATCH) + al16(sizeof(float) * NF) +
               al16(strips_bytes(hop, support, from_spec)) + al16(sizeof(double) * 4 * (kWideThreads / 32)) +
               al16(sizeof(float) * F * MP) + 16 + al16(sizeof(float) * tile_samples);
    }
};

struct TwWide {
    const float2* p;  // [16 k1][32 lanes]
    int l;
    __device__ __forceinline__ float2 operator()(int k1) const { return p[k1 * 32 + l]; }
};

__device__ __forceinline__ void wide_fft_forward(float2* v, int l, TwWide tw, float* scr) {
    w512::fwd_cols(v, tw);
    __syncwarp();
    w512::scr_store_cols(v, l, scr, false);
    __syncwarp();
    w512::scr_load_rows(v, l, scr, false);
    __syncwarp();
    w512::scr_store_cols(v, l, scr, true);
    __syncwarp();
    w512::scr_load_rows(v, l, scr, true);
    w512::fwd_rows_local(v);
    float2 other[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        other[m].x = __shfl_xor_sync(0xffffffffu, v[m].x, 1);
        other[m].y = __shfl_xor_sync(0xffffffffu, v[m].y, 1);
    }
    w512::fwd_rows_combine(v, l, other);
}
__device__ __forceinline__ void wide_fft_inverse(float2* v, int l, TwWide tw, float* scr) {
    float2 other[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        other[m].x = __shfl_xor_sync(0xffffffffu, v[m].x, 1);
        other[m].y = __shfl_xor_sync(0xffffffffu, v[m].y, 1);
    }
    w512::inv_rows_combine(v, l, other);
    w512::inv_rows_local(v);
    __syncwarp();
    w512::scr_store_rows(v, l, scr, false);
    __syncwarp();
    w512::scr_load_cols(v, l, scr, false);
    __syncwarp();
    w512::scr_store_rows(v, l, scr, true);
    __syncwarp();
    w512::scr_load_cols(v, l, scr, true);
    w512::inv_cols(v, tw);
}
__device__ __forceinline__ void wide_exchange8(const float2* send, float2* recv, int src) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        recv[j].x = __shfl_sync(0xffffffffu, send[j].x, src);
        recv[j].y = __shfl_sync(0xffffffffu, send[j].y, src);
    }
}

template <int MODE, bool FROM_SPEC, bool RECT>  // RECT: all-ones window over the whole frame (no window math)
__global__ void __launch_bounds__(kWideThreads, 1)
explain_w512_kernel(PlanDev P, Tiling TL, const float* __restrict__ wav, int64_t wav_stride,
                    const float2* __restrict__ X, int64_t sb, int64_t st, int64_t sf,
                    const float* __restrict__ mask, int Fm, int Tm, int drop,
                    float* __restrict__ rel, float* __restrict__ irr, double* __restrict__ stats) {
    using C = WideCfg;
    constexpr int UNITS = C::UNITS, FT = C::FT, MP = C::MP, F = C::F, NF = C::NF, NT = kWideThreads;
    const int support = P.whi - P.wlo, lb = P.hop + support, strip = C::strip(P.hop, support);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Carver cv{smem_raw};
    float2* tw_s = cv.take<float2>(512);
    float* scratch = cv.take<float>(UNITS * w512::SCRATCH);
    float* win_s = cv.take<float>(NF);
    float2* pb = reinterpret_cast<float2*>(cv.take<unsigned char>(C::strips_bytes(P.hop, support, FROM_SPEC)));
    float* seg = reinterpret_cast<float*>(pb);
    double* red = cv.take<double>(4 * (NT / 32));
    float* mask_s = cv.take<float>(F * MP);
    uint64_t* bar = cv.take<uint64_t>(1);
    float* env_s = cv.take<float>(TL.hops_per_tile * P.hop);

    const int tid = threadIdx.x, b = blockIdx.y;
    const TileGeom g = tile_geom<NF>(P, TL, blockIdx.x);
    const int S = g.s1 - g.s0;
    stage_env<NT>(env_s, P.inv_env + g.s0, S);  // needed only by the epilogue
    bool bulk = false;
    int shift = 0;
    if (!FROM_SPEC)
        shift = stage_segment<NT>(seg, (FT - 1) * P.hop + NF, wav + (size_t)b * wav_stride, g.t_lo * P.hop - NF / 2,
                                  P.n_in, bar, bulk);
    {   // mask tile [F][FT]: every load of the tile (17 per thread) is issued before the first use, so the
        // tile costs one memory latency instead of one per loop trip
        const float* mrow = mask + (size_t)b * Fm * Tm;
        constexpr int kTrips = (F * FT + NT - 1) / NT;
        float mreg[kTrips];
        const float2 twv = P.tw[tid];
        const float wv = P.window[tid];
#pragma unroll
        for (int j = 0; j < kTrips; ++j) {
            const int e = tid + j * NT;
            const int f = e / FT, c = e % FT;
            const int t = g.t_lo + c;
            mreg[j] = (e < F * FT && f < Fm && t < Tm && t <= g.t_hi) ? __ldg(mrow + (size_t)f * Tm + t) : 0.0f;
        }
        // plan table is [32][16] (exp(-2 pi i a b / 512), symmetric in a, b): transpose to [16 k1][32 lanes]
        tw_s[(tid & 15) * 32 + (tid >> 4)] = twv;
        win_s[tid] = wv;
#pragma unroll
        for (int j = 0; j < kTrips; ++j) {
            const int e = tid + j * NT;
            if (e < F * FT) mask_s[(e / FT) * MP + (e % FT)] = mreg[j];
        }
    }
    __syncthreads();

    const int u = tid >> 5, l = tid & 31;
    float* my = scratch + u * w512::SCRATCH;
    const TwWide tw{tw_s, l};
    const int fa = g.t_lo + 2 * u;
    const float* wl = win_s + l;
    const int prt = w512::partner_row(l);

    float2 v[16];
    float2 xa[9], xb[9];
    if (FROM_SPEC) {
        const bool va = fa <= g.t_hi, vb = fa + 1 <= g.t_hi;
        const float2* xa_p = X + (size_t)b * sb + (size_t)fa * st;
        const float2* xb_p = xa_p + st;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const int bin = w512::bin_of(l, i);
            const float2 z = make_float2(0.f, 0.f);
            xa[i] = (va && bin >= 0) ? __ldg(xa_p + (size_t)bin * sf) : z;
            xb[i] = (vb && bin >= 0) ? __ldg(xb_p + (size_t)bin * sf) : z;
        }
    } else {
        stage_wait(bar, bulk);
        const float* sa = seg + shift + (2 * u) * P.hop + l;
        const float* sbp = sa + P.hop;
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
            const float w = RECT ? 1.0f : wl[n1 * 32];
            v[n1] = make_float2(sa[n1 * 32] * w, sbp[n1 * 32] * w);
        }
        __syncthreads();  // the segment's memory becomes the strips
        wide_fft_forward(v, l, tw, my);
        float2 send[8], recv[8];
        w512::split_pre(v, send);
        wide_exchange8(send, recv, prt);
        w512::split_post(v, l, recv, xa, xb);
    }

    float2* pbu = pb + u * strip;
    const int c0 = l - P.wlo, ovl = support - P.hop;
    if (ovl < 0)
        for (int k = support + l; k < P.hop; k += 32) pbu[k] = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        const int t = fa + half;
        const bool valid = t <= g.t_hi;
        const int col = 2 * u + half;
        {
            float2 yr[9], yi[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const int bin = w512::bin_of(l, i);
                const bool live = bin >= 0 && valid;
                const float m = live ? mask_s[bin * MP + col] : 0.0f;
                float gr, gi;
                mask_gains<MODE>(xa[i], m, gr, gi);
                yr[i] = live ? make_float2(xa[i].x * gr, xa[i].y * gr) : make_float2(0.f, 0.f);
                const bool keep_i = live && !(drop && (bin >= Fm || t >= Tm));
                yi[i] = keep_i ? make_float2(xa[i].x * gi, xa[i].y * gi) : make_float2(0.f, 0.f);
            }
            float2 send[8], recv[8];
            w512::merge_pre(v, l, yr, yi, send);
            wide_exchange8(send, recv, prt);
            w512::merge_post(v, l, recv);
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) xa[i] = xb[i];
        wide_fft_inverse(v, l, tw, my);
        float2* dst = pbu + half * P.hop;
        const int keep = half ? ovl : 0;
        if (RECT) {  // support == n_fft, w == 1: every sample lands, only the overlap test remains
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) {
                const int k = n1 * 32 + l;
                float2 o = v[n1];
                if (k < keep) {
                    const float2 prev = dst[k];
                    o.x += prev.x;
                    o.y += prev.y;
                }
                dst[k] = o;
            }
        } else {
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) {
                const int k = n1 * 32 + c0;
                if ((unsigned)k < (unsigned)support) {
                    const float w = wl[n1 * 32];
                    float2 o = k < keep ? dst[k] : make_float2(0.f, 0.f);
                    o.x = fmaf(v[n1].x, w, o.x);
                    o.y = fmaf(v[n1].y, w, o.y);
                    dst[k] = o;
                }
            }
        }
        __syncwarp();  // frame a's strip stores are visible to the unit before frame b's read-modify-write
    }
    cp_async_wait_all();
    __syncthreads();

    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    float* rrow = rel + (size_t)b * P.n_out + g.s0;
    float* irow = irr + (size_t)b * P.n_out + g.s0;
    const int two_hop = 2 * P.hop, x0 = g.p0 - (g.t_lo * P.hop + P.wlo);
    const float inv_two_hop = 1.0f / (float)two_hop;
    for (int q = tid; q < S; q += NT) {
        const float e = env_s[q];
        const float2 o = gather2<UNITS>(pb, strip, lb, two_hop, inv_two_hop, x0 + q);
        const float yr = o.x * e, yi = o.y * e;
        rrow[q] = yr;
        irow[q] = yi;
        acc[0] += (double)yr;
        acc[1] += (double)yr * (double)yr;
        acc[2] += (double)yi;
        acc[3] += (double)yi * (double)yi;
    }
    if (stats != nullptr) {
        block_sum<4, NT>(acc, red);
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
// fused explain, n_fft = 512, persistent + software-pipelined (the production kernel of the hot path)
//
// One CTA per SM (512 threads = 16 wide units) walks the tile list.  While tile i is being transformed the
// loads of tile i+1 are in flight and none of them passes through registers:
//   * waveform segment: one bulk-async copy on an mbarrier (the buffer is free as soon as every thread holds
//     its samples), reflect-padded clip edges by plain loads;
//   * mask tile [257][32]: 4-byte cp.async with zero fill straight into the transposed shared-memory tile of
//     the OTHER parity (two tiles), so the mask is never staged in registers;
//   * plan tables once per CTA.
// The overlap-add gather handles four consecutive samples per step (two 128-bit strip loads per covering
// strip, 128-bit output stores, one index computation per group), the reciprocal envelope of those groups is
// fetched into registers before the barrier, and the lane-pair radix-2 of the wide FFT is the lean form
// (fft_core.cuh: twiddle on the odd lane, one FMA per component) instead of six selects per value.
// Requires hop % 4 == 0, n_out % 4 == 0 and 16-byte aligned output rows (the launcher falls back to
// explain_w512_kernel otherwise).
// ------------------------------------------------------------------------------------------------
struct PWideCfg {
    static constexpr int UNITS = 16, FT = 32, MP = FT + 1, F = 257, NF = 512;
    static __host__ __device__ int strip(int hop, int support) { return (hop + support + 4 + 3) & ~3; }
    static __host__ __device__ size_t seg_floats(int hop) { return (size_t)(FT - 1) * hop + NF + 8; }
    static size_t bytes(int hop, int support) {
        return al16(sizeof(float2) * 512) + al16(sizeof(float) * UNITS * w512::SCRATCH) + al16(sizeof(float) * NF) +
               al16(sizeof(float2) * UNITS * strip(hop, support)) + al16(sizeof(float) * seg_floats(hop)) +
               al16(sizeof(double) * 4 * (kWideThreads / 32)) + 2 * al16(sizeof(float) * F * MP) + 16;
    }
};

__device__ __forceinline__ void lean_fft_forward(float2* v, int l, TwWide tw, float* scr) {
    w512::fwd_cols(v, tw);
    __syncwarp();
    w512::scr_store_cols(v, l, scr, false);
    __syncwarp();
    w512::scr_load_rows(v, l, scr, false);
    __syncwarp();
    w512::scr_store_cols(v, l, scr, true);
    __syncwarp();
    w512::scr_load_rows(v, l, scr, true);
    w512::fwd_rows_local(v);
    w512::fwd_rows_tw(v, l);
    float2 other[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        other[m].x = __shfl_xor_sync(0xffffffffu, v[m].x, 1);
        other[m].y = __shfl_xor_sync(0xffffffffu, v[m].y, 1);
    }
    w512::rows_bfly(v, l, other);
}
__device__ __forceinline__ void lean_fft_inverse(float2* v, int l, TwWide tw, float* scr) {
    float2 other[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        other[m].x = __shfl_xor_sync(0xffffffffu, v[m].x, 1);
        other[m].y = __shfl_xor_sync(0xffffffffu, v[m].y, 1);
    }
    w512::rows_bfly(v, l, other);
    w512::inv_rows_tw(v, l);
    w512::inv_rows_local(v);
    __syncwarp();
    w512::scr_store_rows(v, l, scr, false);
    __syncwarp();
    w512::scr_load_cols(v, l, scr, false);
    __syncwarp();
    w512::scr_store_rows(v, l, scr, true);
    __syncwarp();
    w512::scr_load_cols(v, l, scr, true);
    w512::inv_cols(v, tw);
}

// Register cap of the persistent explain kernel.  128 would be free (512 threads, one CTA per SM) and is 2.5 %
// faster for the kernel alone (84.2 vs 86.3 us), but at 96 the SM keeps 16 K registers for two CTAs of the
// memory-bound normaliser, which then runs NEXT TO the issue-bound explain kernel of the following batch instead of
// after it: the pipelined step drops from 90 to ~81 us (measured, bench.py).
#ifndef ADV_EXPLAIN_MAXREG
#define ADV_EXPLAIN_MAXREG 96
#endif
template <int MODE, bool RECT>
__global__ void __maxnreg__(ADV_EXPLAIN_MAXREG)
explain_p512_kernel(PlanDev P, Tiling TL, int total_tiles, const float* __restrict__ wav, int64_t wav_stride,
                    const float* __restrict__ mask, int Fm, int Tm, float* __restrict__ rel, float* __restrict__ irr,
                    double* __restrict__ stats) {
    using C = PWideCfg;
    constexpr int UNITS = C::UNITS, FT = C::FT, MP = C::MP, F = C::F, NF = C::NF, NT = kWideThreads;
    const int wlo = P.wlo & ~3;  // strip origin rounded down to a multiple of 4 (extra taps are exact zeros)
    const int support = P.whi - wlo, lb = P.hop + support, strip = C::strip(P.hop, P.whi - P.wlo);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Carver cv{smem_raw};
    float2* tw_s = cv.take<float2>(512);
    float* scratch = cv.take<float>(UNITS * w512::SCRATCH);
    float* win_s = cv.take<float>(NF);
    float2* pb = cv.take<float2>(UNITS * strip);
    float* seg = cv.take<float>(C::seg_floats(P.hop));
    double* red = cv.take<double>(4 * (NT / 32));
    float* mask_all = cv.take<float>(2 * ((F * MP + 3) & ~3));
    uint64_t* bar = cv.take<uint64_t>(1);
    constexpr int kMaskTile = (F * MP + 3) & ~3;

    const int tid = threadIdx.x, u = tid >> 5, l = tid & 31;
    const int seglen = (FT - 1) * P.hop + NF;
    int tile = blockIdx.x;
    int b = tile / TL.tiles;
    TileGeom g = tile_geom<NF>(P, TL, tile - b * TL.tiles);

    // every load of a tile that can be requested ahead of time: segment (bulk + plain edges), mask tile
    auto request_tile = [&](int bb, const TileGeom& gg, float* mask_dst) {
        stage_segment_async<NT>(seg, seglen, wav + (size_t)bb * wav_stride, gg.t_lo * P.hop - NF / 2, P.n_in, bar);
        // thread -> fixed column c = tid & 31 and rows f = (tid >> 5) + 16 j: source and destination advance by a
        // constant per step, the column test is loop-invariant
        const int c = tid & 31, f0 = tid >> 5;
        const int t = gg.t_lo + c;
        const bool col_ok = t < Tm && t <= gg.t_hi;
        const float* mrow = mask + (size_t)bb * Fm * Tm;
        const float* src = mrow + (size_t)f0 * Tm + (col_ok ? t : 0);
        uint32_t dst = smem_u32(mask_dst + f0 * MP + c);
        constexpr int kRowsPerStep = NT / 32, kTrips = (F + kRowsPerStep - 1) / kRowsPerStep;
#pragma unroll
        for (int j = 0; j < kTrips; ++j) {
            const int f = f0 + j * kRowsPerStep;
            if (f < F) {
                const bool ok = col_ok && f < Fm;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(ok ? src : mrow), "r"(ok ? 4 : 0)
                             : "memory");
            }
            src += (size_t)kRowsPerStep * Tm;
            dst += kRowsPerStep * MP * 4;
        }
        cp_async_commit();
    };

    if (tid == 0) mbar_init(bar, 1);
    {   // plan table is [32][16] (exp(-2 pi i a b / 512), symmetric in a, b): transpose to [16 k1][32 lanes]
        tw_s[(tid & 15) * 32 + (tid >> 4)] = P.tw[tid];
        win_s[tid] = P.window[tid];
    }
    pdl_launch_dependents();
    __syncthreads();
    pdl_wait();  // (everything above touches plan constants only)
    request_tile(b, g, mask_all);
    const int shift0 = (g.t_lo * P.hop - NF / 2) & 3;
    int shift = shift0;

    float* my = scratch + u * w512::SCRATCH;
    const TwWide tw{tw_s, l};
    const float* wl = win_s + l;
    const int prt = w512::partner_row(l);
    float2* pbu = pb + u * strip;
    const int c0 = l - wlo, ovl = support - P.hop, two_hop = 2 * P.hop;
    const float inv_two_hop = 1.0f / (float)two_hop;

    for (uint32_t it = 0;; ++it) {
        const float* mask_s = mask_all + (it & 1) * kMaskTile;
        cp_async_wait_all();
        __syncthreads();  // mask tile + plain part of the segment visible; the previous epilogue is over
        mbar_wait(bar, it & 1);
        float2 v[16];
        {
            const float* sa = seg + shift + (2 * u) * P.hop + l;
            const float* sbp = sa + P.hop;
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) {
                const float w = RECT ? 1.0f : wl[n1 * 32];
                v[n1] = make_float2(sa[n1 * 32] * w, sbp[n1 * 32] * w);
            }
        }
        const int cur_b = b, cur_tile = tile - b * TL.tiles;
        const TileGeom cg = g;
        const int next = tile + gridDim.x;
        __threadfence_block();  // the sample loads have returned: a sync alone does not wait for queued shared-memory loads
        __syncthreads();  // every thread holds its samples: the segment buffer is free
        if (next < total_tiles) {
            b = next / TL.tiles;
            g = tile_geom<NF>(P, TL, next - b * TL.tiles);
            request_tile(b, g, mask_all + ((it + 1) & 1) * kMaskTile);
            shift = (g.t_lo * P.hop - NF / 2) & 3;
        }

        const int fa = cg.t_lo + 2 * u;
        if (fa <= cg.t_hi) {  // (warp-uniform) units past the tile's last frame have nothing to transform
        float2 xa[9], xb[9];
        lean_fft_forward(v, l, tw, my);
        {
            float2 send[8], recv[8];
            w512::split_pre(v, send);
            wide_exchange8(send, recv, prt);
            w512::split_post(v, l, recv, xa, xb);
        }
        if (ovl < 0)
            for (int k = support + l; k < P.hop; k += 32) pbu[k] = make_float2(0.f, 0.f);
        // fully unrolled: a rolled loop has to rotate frame b's spectrum into frame a's registers (5 % of the
        // kernel's instructions were moves); the body is ~1.8 k SASS instructions, the kernel stays under 100 KB
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const bool valid = fa + half <= cg.t_hi;
            const int col = 2 * u + half;
            {
                float2 yr[9], yi[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) {
                    const int bin = w512::bin_of(l, i);
                    const bool live = bin >= 0 && valid;
                    const float m = live ? mask_s[bin * MP + col] : 0.0f;
                    float gr, gi;
                    mask_gains<MODE>(xa[i], m, gr, gi);
                    yr[i] = live ? make_float2(xa[i].x * gr, xa[i].y * gr) : make_float2(0.f, 0.f);
                    yi[i] = live ? make_float2(xa[i].x * gi, xa[i].y * gi) : make_float2(0.f, 0.f);
                }
                float2 send[8], recv[8];
                w512::merge_pre(v, l, yr, yi, send);
                wide_exchange8(send, recv, prt);
                w512::merge_post(v, l, recv);
            }
#pragma unroll
            for (int i = 0; i < 9; ++i) xa[i] = xb[i];
            lean_fft_inverse(v, l, tw, my);
            float2* dst = pbu + half * P.hop;
            const int keep = half ? ovl : 0;
            if (RECT) {
#pragma unroll
                for (int n1 = 0; n1 < 16; ++n1) {
                    const int k = n1 * 32 + l;
                    float2 o = v[n1];
                    if (k < keep) {
                        const float2 prev = dst[k];
                        o.x += prev.x;
                        o.y += prev.y;
                    }
                    dst[k] = o;
                }
            } else {
#pragma unroll
                for (int n1 = 0; n1 < 16; ++n1) {
                    const int k = n1 * 32 + c0;
                    if ((unsigned)k < (unsigned)support) {
                        const float w = wl[n1 * 32];
                        float2 o = k < keep ? dst[k] : make_float2(0.f, 0.f);
                        o.x = fmaf(v[n1].x, w, o.x);
                        o.y = fmaf(v[n1].y, w, o.y);
                        dst[k] = o;
                    }
                }
            }
            __syncwarp();  // frame a's strip stores are visible to the unit before frame b's read-modify-write
        }
        } else {
            // the gather may still touch this strip (frames clipped at the end of the clip): it must read zeros
            for (int k = l; k < lb; k += 32) pbu[k] = make_float2(0.f, 0.f);
        }

        // reciprocal envelope of this thread's first groups: requested before the barrier
        const int S = cg.s1 - cg.s0;
        const float* erow = P.inv_env + cg.s0;
        constexpr int kEnvRegs = 3;
        float4 e4[kEnvRegs];
#pragma unroll
        for (int j = 0; j < kEnvRegs; ++j) {
            const int q = (tid + j * NT) * 4;
            e4[j] = q + 4 <= S ? __ldg(reinterpret_cast<const float4*>(erow + q)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();

        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        float* rrow = rel + (size_t)cur_b * P.n_out + cg.s0;
        float* irow = irr + (size_t)cur_b * P.n_out + cg.s0;
        const int x0 = cg.p0 - (cg.t_lo * P.hop + wlo);
        auto group = [&](int q, float4 e) {
            const int x = x0 + q;
            int uu = min(UNITS - 1, (int)(((float)x + 0.5f) * inv_two_hop));
            int k = x - uu * two_hop;
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f), ir = r;
            while (uu >= 0 && k < lb) {
                const float4* p = reinterpret_cast<const float4*>(pb + uu * strip + k);
                const float4 a = p[0], c = p[1];
                r.x += a.x; ir.x += a.y; r.y += a.z; ir.y += a.w;
                r.z += c.x; ir.z += c.y; r.w += c.z; ir.w += c.w;
                --uu;
                k += two_hop;
            }
            r.x *= e.x; r.y *= e.y; r.z *= e.z; r.w *= e.w;
            ir.x *= e.x; ir.y *= e.y; ir.z *= e.z; ir.w *= e.w;
            *reinterpret_cast<float4*>(rrow + q) = r;
            *reinterpret_cast<float4*>(irow + q) = ir;
            acc[0] += (double)((r.x + r.y) + (r.z + r.w));
            acc[1] += (double)(fmaf(r.x, r.x, r.y * r.y) + fmaf(r.z, r.z, r.w * r.w));
            acc[2] += (double)((ir.x + ir.y) + (ir.z + ir.w));
            acc[3] += (double)(fmaf(ir.x, ir.x, ir.y * ir.y) + fmaf(ir.z, ir.z, ir.w * ir.w));
        };
#pragma unroll
        for (int j = 0; j < kEnvRegs; ++j) {
            const int q = (tid + j * NT) * 4;
            if (q + 4 <= S) group(q, e4[j]);
        }
        for (int q = (tid + kEnvRegs * NT) * 4; q + 4 <= S; q += NT * 4)  // (tiles longer than 6144 samples)
            group(q, __ldg(reinterpret_cast<const float4*>(erow + q)));
        if (stats != nullptr) {
            block_sum<4, NT>(acc, red);
            if (tid == 0) {
                double* srow = stats + ((size_t)cur_b * TL.tiles + cur_tile) * 4;
                srow[0] = acc[0];
                srow[1] = acc[1];
                srow[2] = acc[2];
                srow[3] = acc[3];
            }
        }
        if (next >= total_tiles) break;
        tile = next;
    }
}

// ------------------------------------------------------------------------------------------------
// STFT, n_fft = 512, wide units + warp-autonomous (one warp = one complex FFT = two frames per item)
//
// Same scheme as stft_w_kernel (every warp owns its frames, its slice of the waveform, its mbarrier and its place
// in the item list; the next slice is requested by one bulk-async copy as soon as the warp holds the current
// samples), but with the 32-lane x 16-value FFT of the fused kernel: 64 registers per thread instead of 128, so
// FOUR 256-thread CTAs share an SM (32 warps instead of 16) and the sqrt / atan2 / store tail of one warp hides
// behind the butterflies of the others.  Lane (r, h) ends up with bins bin0 + 16 i of both frames: every store
// instruction of the warp writes two full 128-byte runs of the row.  (Four-frame items - two FFTs per staged slice -
// were measured: 32.7 vs 31.8 us, 145 vs 130 us at 256 clips; the larger slices cost a resident CTA.)
// ------------------------------------------------------------------------------------------------
struct WWCfg {
    static constexpr int WARPS = kThreads / 32, NF = 512, F = 257;
    static __host__ __device__ size_t seg_floats(int hop) { return ((size_t)hop + NF + 8 + 3) & ~size_t(3); }  // 16-byte pitch
    static size_t bytes(int hop) {
        return al16(sizeof(float2) * 512) + al16(sizeof(float) * WARPS * w512::SCRATCH) + al16(sizeof(float) * NF) +
               al16(sizeof(float) * WARPS * seg_floats(hop)) + al16(sizeof(uint64_t) * WARPS);
    }
};

template <bool MAG, bool PHASE, bool RECT>
__global__ void __launch_bounds__(kThreads, 4)
stft_ww_kernel(PlanDev P, const float* __restrict__ wav, int64_t wav_stride, int total_items, int items_per_clip,
               float2* __restrict__ X, float* __restrict__ mag, float* __restrict__ phase) {
    using C = WWCfg;
    constexpr int F = C::F, NF = C::NF;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Carver cv{smem_raw};
    float2* tw_s = cv.take<float2>(512);
    float* scratch = cv.take<float>(C::WARPS * w512::SCRATCH);
    float* win_s = cv.take<float>(NF);
    float* seg_all = cv.take<float>(C::WARPS * C::seg_floats(P.hop));
    uint64_t* bars = cv.take<uint64_t>(C::WARPS);

    const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
    float* seg = seg_all + (size_t)w * C::seg_floats(P.hop);
    uint64_t* bar = bars + w;
    const int seglen = P.hop + NF;
    const int stride = gridDim.x * C::WARPS;
    int item = blockIdx.x * C::WARPS + w;

    if (l == 0) mbar_init(bar, 1);
    for (int i = tid; i < 512; i += kThreads) tw_s[(i & 15) * 32 + (i >> 4)] = P.tw[i];  // [32][16] -> [16 k1][32 lanes]
    for (int i = tid; i < NF; i += kThreads) win_s[i] = P.window[i];
    pdl_launch_dependents();
    __syncthreads();  // tables staged, barriers initialised
    pdl_wait();
    if (item >= total_items) return;

    float* my = scratch + w * w512::SCRATCH;
    const TwWide tw{tw_s, l};
    const int prt = w512::partner_row(l);
    const int bin0 = (l & 1) ? 128 + ((16 - (l >> 1)) & 15) : (l >> 1);  // w512::bin_of(l, i) = bin0 + 16 i, i < 8
    int b = item / items_per_clip, t0 = (item - b * items_per_clip) * 2;
    int shift = stage_segment_async<32>(seg, seglen, wav + (size_t)b * wav_stride, t0 * P.hop - NF / 2, P.n_in, bar, l);

    for (uint32_t it = 0;; ++it) {
        __syncwarp();  // plain-load part of the slice visible to the warp
        mbar_wait(bar, it & 1);
        float2 v[16];
        {
            const float* sa = seg + shift + l;
            const float* sb = sa + P.hop;
            const float* wl = win_s + l;
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) {
                const float ww = RECT ? 1.0f : wl[n1 * 32];
                v[n1] = make_float2(sa[n1 * 32] * ww, sb[n1 * 32] * ww);
            }
        }
        const int cur_b = b, fa = t0;
        const int next = item + stride;
        __threadfence_block();  // the sample loads have returned: a sync alone does not wait for queued shared-memory loads
        __syncwarp();  // every lane holds its samples: the slice buffer is free for the next item
        if (next < total_items) {
            b = next / items_per_clip;
            t0 = (next - b * items_per_clip) * 2;
            shift = stage_segment_async<32>(seg, seglen, wav + (size_t)b * wav_stride, t0 * P.hop - NF / 2, P.n_in, bar, l);
        }
        lean_fft_forward(v, l, tw, my);
        float2 xa[9], xb[9];
        {
            float2 send[8], recv[8];
            w512::split_pre(v, send);
            wide_exchange8(send, recv, prt);
            w512::split_post(v, l, recv, xa, xb);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int t = fa + half;
            if (t >= P.T) continue;
            const float2* x = half ? xb : xa;
            const size_t row = ((size_t)cur_b * P.T + t) * F + bin0;
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                if (i == 8 && l != 1) continue;  // slot 8 is the Nyquist bin (256 = bin0 + 128 on lane 1)
                X[row + 16 * i] = x[i];
                if (MAG) mag[row + 16 * i] = fast_abs2(x[i]);
                if (PHASE) phase[row + 16 * i] = fast_atan2f(x[i].y, x[i].x);
            }
        }
        if (next >= total_items) break;
        item = next;
    }
}

// ------------------------------------------------------------------------------------------------
// iSTFT, n_fft = 512, wide units (one warp per complex FFT = two frames), persistent
//
// The narrow-unit inverse (istft_p_kernel) keeps 32 complex values per lane: 128 registers, 16 warps per SM,
// and it is instruction-issue / latency bound far below the HBM roofline.  Here a whole warp transforms the
// frame pair (16 values per lane), so a 512-thread CTA needs <= 64 registers per thread and TWO of them share
// an SM (32 warps): while one CTA waits for its spectrum rows or sits in the gather, the other runs FFTs.
//   * rows: lane (r, h) reads the 8 (+ Nyquist) one-sided bins of w512::bin_of - per load instruction a warp
//     covers two full 128-byte runs of the row;
//   * both frames of the unit go through ONE complex inverse FFT (merge = the explain kernel's);
//   * HS > 0 (rectangular full window, hop = 32 * HS): sample k = 32 j + l of frame b lands on sample
//     32 (j + HS) + l of the unit's strip, i.e. in the SAME lane as frame a's sample - the unit's two frames are
//     overlap-added in registers and the strip is written once (16 + HS stores, no read-modify-write);
//   * gather: four consecutive samples per step as in istft_p_kernel; the reciprocal envelope comes from L2.
// Requires contiguous rows (bin stride 1), hop % 4 == 0, n_out % 4 == 0, 16-byte aligned output rows.
// ------------------------------------------------------------------------------------------------
struct IWCfg {
    static constexpr int UNITS = 16, NF = 512;
    static __host__ __device__ int strip(int hop, int support) { return (hop + support + 4 + 3) & ~3; }
    static size_t bytes(int hop, int support) {
        return al16(sizeof(float2) * 512) + al16(sizeof(float) * UNITS * w512::SCRATCH) + al16(sizeof(float) * NF) +
               al16(sizeof(float) * UNITS * strip(hop, support)) + al16(sizeof(double) * 2 * (kWideThreads / 32));
    }
};

template <bool RECT, int HS>
__global__ void __launch_bounds__(kWideThreads, 2)
istft_w512_kernel(PlanDev P, Tiling TL, int total_tiles, const float2* __restrict__ X, int64_t sb, int64_t st,
                  float* __restrict__ out, double* __restrict__ stats) {
    using C = IWCfg;
    constexpr int UNITS = C::UNITS, NF = C::NF, NT = kWideThreads;
    // HS > 0: hop, support and the strip pitch are compile-time constants (index arithmetic of the gather folds)
    constexpr bool FIX = RECT && HS > 0;
    const int hop = FIX ? 32 * HS : P.hop;
    const int wlo = FIX ? 0 : P.wlo & ~3;  // strip origin rounded down to a multiple of 4 (extra taps are exact zeros)
    const int support = FIX ? NF : P.whi - wlo, lb = hop + support;
    const int strip = FIX ? C::strip(32 * HS, NF) : C::strip(P.hop, P.whi - P.wlo);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Carver cv{smem_raw};
    float2* tw_s = cv.take<float2>(512);
    float* scratch = cv.take<float>(UNITS * w512::SCRATCH);
    float* win_s = cv.take<float>(NF);
    float* pb = cv.take<float>(UNITS * strip);
    double* red = cv.take<double>(2 * (NT / 32));

    const int tid = threadIdx.x, u = tid >> 5, l = tid & 31;
    tw_s[(tid & 15) * 32 + (tid >> 4)] = P.tw[tid];  // plan table [32][16] -> [16 k1][32 lanes]
    win_s[tid] = P.window[tid];
    pdl_launch_dependents();

    // lane (r, h): slots 0..7 are bins bin0 + 16 i (w512::bin_of), slot 8 is the Nyquist bin on lane 1 only
    const int bin0 = (l & 1) ? 128 + ((16 - (l >> 1)) & 15) : (l >> 1);
    float2 xa[9], xb[9];
    auto load_rows = [&](int t_id) {
        const int bb = t_id / TL.tiles;
        const TileGeom gg = tile_geom<NF>(P, TL, t_id - bb * TL.tiles);
        const int fa = gg.t_lo + 2 * u;
        const float2* xa_p = X + (size_t)bb * sb + (size_t)fa * st + bin0;
        const float2* xb_p = xa_p + st;
        const float2 z = make_float2(0.f, 0.f);
        if (fa + 1 <= gg.t_hi) {  // (warp-uniform) both frames live: the common case
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                xa[i] = __ldg(xa_p + 16 * i);
                xb[i] = __ldg(xb_p + 16 * i);
            }
            xa[8] = l == 1 ? __ldg(xa_p + 128) : z;
            xb[8] = l == 1 ? __ldg(xb_p + 128) : z;
        } else {
            const bool va = fa <= gg.t_hi;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                xa[i] = va ? __ldg(xa_p + 16 * i) : z;
                xb[i] = z;
            }
            xa[8] = (va && l == 1) ? __ldg(xa_p + 128) : z;
            xb[8] = z;
        }
    };
    // next tile's rows -> L2 (no registers held across the gather): lane l touches line l of each row
    auto prefetch_rows = [&](int t_id) {
        const int bb = t_id / TL.tiles;
        const TileGeom gg = tile_geom<NF>(P, TL, t_id - bb * TL.tiles);
        const int fa = gg.t_lo + 2 * u;
        if (fa + 1 <= gg.t_hi && l * 16 < 257) {
            const float2* xa_p = X + (size_t)bb * sb + (size_t)fa * st + l * 16;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(xa_p));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(xa_p + st));
        }
    };
    int tile = blockIdx.x;
    __syncthreads();
    pdl_wait();  // (everything above touches plan constants only)

    float* my = scratch + u * w512::SCRATCH;
    const TwWide tw{tw_s, l};
    const float* wl = win_s + l;
    const int prt = w512::partner_row(l);
    float* pbu = pb + u * strip;
    const int c0 = l - wlo, ovl = support - hop, two_hop = 2 * hop;
    const float inv_two_hop = 1.0f / (float)two_hop;

    for (;;) {
        const int b = tile / TL.tiles, cur_tile = tile - b * TL.tiles;
        const TileGeom g = tile_geom<NF>(P, TL, cur_tile);
        const int next = tile + gridDim.x;
        if (g.t_lo + 2 * u <= g.t_hi) {  // (warp-uniform) units past the tile's last frame have nothing to transform
            load_rows(tile);
            float2 v[16];
            {
                float2 send[8], recv[8];
                w512::merge_pre(v, l, xa, xb, send);
                wide_exchange8(send, recv, prt);
                w512::merge_post(v, l, recv);
            }
            lean_fft_inverse(v, l, tw, my);
            // private strip of the unit: frame a at [0, support), frame b at [hop, hop + support)
            // (the previous tile's gather finished reading the strips: barrier at the loop end)
            if (RECT && HS > 0) {
#pragma unroll
                for (int j = 0; j < 16 + HS; ++j) {
                    float o = j < 16 ? v[j < 16 ? j : 0].x : 0.0f;
                    if (j >= HS) o += v[j >= HS ? j - HS : 0].y;
                    pbu[j * 32 + l] = o;
                }
            } else {
#pragma unroll
                for (int n1 = 0; n1 < 16; ++n1) {
                    const int k = n1 * 32 + c0;
                    if ((unsigned)k < (unsigned)support) pbu[k] = RECT ? v[n1].x : v[n1].x * wl[n1 * 32];
                }
                if (ovl < 0)  // degenerate (hop > support, only legal for single-frame plans): clear the gap
                    for (int k = support + l; k < hop; k += 32) pbu[k] = 0.0f;
                __syncwarp();
#pragma unroll
                for (int n1 = 0; n1 < 16; ++n1) {
                    const int k = n1 * 32 + c0;
                    if ((unsigned)k < (unsigned)support) {
                        const float old = k < ovl ? pbu[hop + k] : 0.0f;
                        pbu[hop + k] = RECT ? v[n1].y + old : fmaf(v[n1].y, wl[n1 * 32], old);
                    }
                }
            }
        } else {
            // the gather may still touch this strip (frames clipped at the end of the clip): it must read zeros
            for (int k = l; k < lb; k += 32) pbu[k] = 0.0f;
        }
        // Holding the next rows in registers across the gather (or from the end of the gather across the barrier)
        // was measured: 36 more live registers spill ~600 bytes per thread at the 64-register cap and the kernel
        // takes 47 us instead of 30.  The L2 prefetch costs two instructions and no registers.
        if (next < total_tiles) prefetch_rows(next);
        __syncthreads();

        // gather: 4 samples per step; strips covering offset x are u = x / two_hop, u-1, ... while k < lb
        const int S = g.s1 - g.s0;
        const int x0 = g.p0 - (g.t_lo * hop + wlo);
        float* orow = out + (size_t)b * P.n_out + g.s0;
        const float* erow = P.inv_env + g.s0;
        double acc0 = 0.0, acc1 = 0.0;
        for (int q = tid * 4; q < S; q += NT * 4) {  // (n_out % 4 == 0 and hop % 4 == 0: S % 4 == 0)
            const float4 e = __ldg(reinterpret_cast<const float4*>(erow + q));
            const int x = x0 + q;
            int uu = min(UNITS - 1, (int)(((float)x + 0.5f) * inv_two_hop));  // exact: x < 2^20
            int k = x - uu * two_hop;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            while (uu >= 0 && k < lb) {
                vadd(a, *reinterpret_cast<const float4*>(pb + uu * strip + k));
                --uu;
                k += two_hop;
            }
            float sq;
            const float sm = vmul_stats(a, e, sq);
            *reinterpret_cast<float4*>(orow + q) = a;
            acc0 += (double)sm;
            acc1 += (double)sq;
        }
        if (stats != nullptr) {
            double acc[2] = {acc0, acc1};
            block_sum<2, NT>(acc, red);
            if (tid == 0) {
                double* srow = stats + ((size_t)b * TL.tiles + cur_tile) * 2;
                srow[0] = acc[0];
                srow[1] = acc[1];
            }
        }
        if (next >= total_tiles) break;
        tile = next;
        __syncthreads();  // strips and reduction scratch are free again
    }
}

// ------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------

#ifdef ADV_AB
template <int NF>
static int launch_stft_nf(const adv_plan* p, const float* wav, int64_t wav_stride, int batch, float2* X,
                          float* mag, float* phase, cudaStream_t s) {
    constexpr int FT = Cfg<NF>::FT;
    const size_t smem = Cfg<NF>::stft_bytes(p->d.hop);
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


template <int NF, bool RECT>
static int launch_stft_p(const adv_plan* p, const float* wav, int64_t wav_stride, int batch, float2* X, float* mag,
                         float* phase, cudaStream_t s) {
    constexpr int FT = PCfg<NF>::FT;
    const size_t smem = PCfg<NF>::stft_bytes(p->d.hop);
    const int tiles_per_clip = (p->d.T + FT - 1) / FT;
    const long total = (long)tiles_per_clip * batch;
    if (total > 0x7fffffffL) return ADV_ERR_UNSUPPORTED;
    const int grid = (int)(total < 2L * sm_count() ? total : 2L * sm_count());
    int rc;
#define ADV_LAUNCH_STFT(M, PH)                                                                              \
    do {                                                                                                    \
        if ((rc = set_smem(stft_p_kernel<NF, M, PH, RECT>, smem)) != ADV_OK) return rc;                     \
        stft_p_kernel<NF, M, PH, RECT><<<grid, kThreads, smem, s>>>(p->d, wav, wav_stride, (int)total,      \
                                                                     tiles_per_clip, X, mag, phase);        \
    } while (0)
    if (mag && phase) ADV_LAUNCH_STFT(true, true);
    else if (mag) ADV_LAUNCH_STFT(true, false);
    else if (phase) ADV_LAUNCH_STFT(false, true);
    else ADV_LAUNCH_STFT(false, false);
#undef ADV_LAUNCH_STFT
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

#endif  // ADV_AB
template <int NF, bool RECT>
static int launch_stft_w(const adv_plan* p, const float* wav, int64_t wav_stride, int batch, float2* X, float* mag,
                         float* phase, int flags, cudaStream_t s) {
    using C = WCfg<NF>;
    const bool zero_pad = (flags & ADV_STFT_ZERO_PAD) != 0;
    // (the BULK = true form of the kernel - finished rows staged in shared memory and written by bulk-async stores -
    //  measured slower, 22.3 vs 17.2 us, and is not instantiated)
    const size_t smem = C::stft_bytes(p->d.hop, false);
    const int items_per_clip = (p->d.T + C::FW - 1) / C::FW;
    const long total = (long)items_per_clip * batch;
    if (total > 0x7fffffffL) return ADV_ERR_UNSUPPORTED;
    const long ctas = (total + C::WARPS - 1) / C::WARPS;
    const int grid = (int)(ctas < 2L * sm_count() ? ctas : 2L * sm_count());
    int rc;
    if (zero_pad) {  // the adjoint-of-istft use (training backward): spectrum only
        if (mag || phase) return ADV_ERR_UNSUPPORTED;
        if ((rc = set_smem(stft_w_kernel<NF, false, false, RECT, false, true>, smem)) != ADV_OK) return rc;
        ADV_CUDA_CHECK(launch_pdl(stft_w_kernel<NF, false, false, RECT, false, true>, grid, kThreads, smem, s, p->d, wav,
                                  wav_stride, (int)total, items_per_clip, X, mag, phase));
        ADV_CUDA_CHECK(cudaGetLastError());
        return ADV_OK;
    }
#define ADV_LAUNCH_STFT(M, PH)                                                                              \
    do {                                                                                                    \
        if ((rc = set_smem(stft_w_kernel<NF, M, PH, RECT, false>, smem)) != ADV_OK) return rc;              \
        ADV_CUDA_CHECK(launch_pdl(stft_w_kernel<NF, M, PH, RECT, false>, grid, kThreads, smem, s, p->d, wav, \
                                  wav_stride, (int)total, items_per_clip, X, mag, phase));                  \
    } while (0)
    if (mag && phase) ADV_LAUNCH_STFT(true, true);
    else if (mag) ADV_LAUNCH_STFT(true, false);
    else if (phase) ADV_LAUNCH_STFT(false, true);
    else ADV_LAUNCH_STFT(false, false);
#undef ADV_LAUNCH_STFT
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

template <bool RECT>
static int launch_stft_ww(const adv_plan* p, const float* wav, int64_t wav_stride, int batch, float2* X, float* mag,
                          float* phase, cudaStream_t s) {
    using C = WWCfg;
    const size_t smem = C::bytes(p->d.hop);
    const int items_per_clip = (p->d.T + 1) / 2;
    const long total = (long)items_per_clip * batch;
    if (total > 0x7fffffffL) return ADV_ERR_UNSUPPORTED;
    const long ctas = (total + C::WARPS - 1) / C::WARPS;
    int rc;
#define ADV_LAUNCH_STFT(M, PH)                                                                                   \
    do {                                                                                                         \
        auto kernel = stft_ww_kernel<M, PH, RECT>;                                                               \
        if ((rc = set_smem(kernel, smem)) != ADV_OK) return rc;                                                  \
        const int resident = resident_memo(kernel, kThreads, smem, 4);                                           \
        const long slots = (long)resident * sm_count();                                                          \
        const int grid = (int)(ctas < slots ? ctas : slots);                                                     \
        ADV_CUDA_CHECK(launch_pdl(kernel, grid, kThreads, smem, s, p->d, wav, wav_stride, (int)total,            \
                                  items_per_clip, X, mag, phase));                                               \
    } while (0)
    if (mag && phase) ADV_LAUNCH_STFT(true, true);
    else if (mag) ADV_LAUNCH_STFT(true, false);
    else if (phase) ADV_LAUNCH_STFT(false, true);
    else ADV_LAUNCH_STFT(false, false);
#undef ADV_LAUNCH_STFT
    return ADV_OK;
}

int launch_stft(const adv_plan* p, const float* wav, int64_t wav_stride, int batch, float2* X, float* mag,
                float* phase, int flags, cudaStream_t s) {
    {   // generation 3 (transform3_kernels.cu) takes every call inside its domain; ADV_GEN3=0 disables it
        const int rc3 = launch_stft3(p, wav, wav_stride, batch, X, mag, phase, flags, s);
        if (rc3 != ADV_ERR_UNSUPPORTED) return rc3;
    }
    {   // n_fft 1024, even hop, reflect padding: one frame per warp, support-only slices (transform5_kernels.cu)
        const int rc5 = launch_stft5(p, wav, wav_stride, batch, X, mag, phase, flags, s);
        if (rc5 != ADV_ERR_UNSUPPORTED) return rc5;
    }
#ifdef ADV_AB
    static const char* var = ADV_AB_ENV("ADV_STFT");  // A/B switch: "v2" one tile per CTA, "p" persistent CTA tiles
    const int which = var == nullptr ? 0 : (var[0] == 'v' ? 2 : (var[0] == 'p' ? 1 : 0));
    if (which != 0 && (flags & ADV_STFT_ZERO_PAD)) return ADV_ERR_UNSUPPORTED;  // older kernels: reflect padding only
    if (which == 2)
        return p->d.n_fft == 512 ? launch_stft_nf<512>(p, wav, wav_stride, batch, X, mag, phase, s)
                                 : launch_stft_nf<1024>(p, wav, wav_stride, batch, X, mag, phase, s);
    if (which == 1) {
        if (p->d.n_fft == 512)
            return p->d.rect_full ? launch_stft_p<512, true>(p, wav, wav_stride, batch, X, mag, phase, s)
                                  : launch_stft_p<512, false>(p, wav, wav_stride, batch, X, mag, phase, s);
        return p->d.rect_full ? launch_stft_p<1024, true>(p, wav, wav_stride, batch, X, mag, phase, s)
                              : launch_stft_p<1024, false>(p, wav, wav_stride, batch, X, mag, phase, s);
    }
#endif  // ADV_AB
#ifdef ADV_AB
    // generation 2, n_fft 512 (reachable with ADV_GEN3=0): the wide-unit kernel when magnitude and / or phase are requested
    // (X + |X| + angle 31.5 us vs 36.3 us on the narrow units per 64 x 4 s clips; X only 18.5 vs 17.1 us).
    // ADV_STFT_WW=0 / =all force the narrow / the wide kernel.
    static const char* ww_env = ADV_AB_ENV("ADV_STFT_WW");
    static const bool ww_on = !(ww_env && ww_env[0] == '0'), ww_all = ww_env && ww_env[0] == 'a';
    if (ww_on && (mag || phase || ww_all) && p->d.n_fft == 512 && !(flags & ADV_STFT_ZERO_PAD))
        return p->d.rect_full ? launch_stft_ww<true>(p, wav, wav_stride, batch, X, mag, phase, s)
                              : launch_stft_ww<false>(p, wav, wav_stride, batch, X, mag, phase, s);
    if (p->d.n_fft == 512)
        return p->d.rect_full ? launch_stft_w<512, true>(p, wav, wav_stride, batch, X, mag, phase, flags, s)
                              : launch_stft_w<512, false>(p, wav, wav_stride, batch, X, mag, phase, flags, s);
#else
    if (p->d.n_fft == 512) return ADV_ERR_UNSUPPORTED;  // generation 3 owns n_fft 512 (plans without it cannot exist)
#endif
    // n_fft 1024: the warp-autonomous kernel (a warp per frame, 32 values per lane) - faster than the generation-3
    // half-size transform (transform3_kernels.cu, launch_stft3)
    return p->d.rect_full ? launch_stft_w<1024, true>(p, wav, wav_stride, batch, X, mag, phase, flags, s)
                          : launch_stft_w<1024, false>(p, wav, wav_stride, batch, X, mag, phase, flags, s);
}

template <int NF>
static int launch_istft_nf(const adv_plan* p, const float2* X, int64_t sb, int64_t st, int64_t sf, int batch,
                           float* out, double* stats, cudaStream_t s) {
    const Tiling tl = choose_tiling(p, batch, 2, istft_balanced());
    const size_t smem = Cfg<NF>::istft_bytes(p->d.hop, p->d.whi - p->d.wlo, tl.hops_per_tile * p->d.hop);
    int rc = set_smem(istft_kernel<NF>, smem);
    if (rc != ADV_OK) return rc;
    dim3 grid(tl.tiles, batch);
    istft_kernel<NF><<<grid, kThreads, smem, s>>>(p->d, tl, X, sb, st, sf, out, stats);
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

template <int NF, int VEC>
static int launch_istft_p(const adv_plan* p, const Tiling& tl, const float2* X, int64_t sb, int64_t st, int64_t sf,
                          int batch, float* out, double* stats, cudaStream_t s) {
    const size_t smem = IPCfg<NF>::bytes(p->d.hop, p->d.whi - p->d.wlo, tl.hops_per_tile * p->d.hop);
    const long total = (long)tl.tiles * batch;
    if (total > 0x7fffffffL) return ADV_ERR_UNSUPPORTED;
    const int grid = (int)(total < 2L * sm_count() ? total : 2L * sm_count());
    int rc;
    if (sf == 1) {
        if ((rc = set_smem(istft_p_kernel<NF, VEC, true>, smem)) != ADV_OK) return rc;
        istft_p_kernel<NF, VEC, true><<<grid, kThreads, smem, s>>>(p->d, tl, (int)total, X, sb, st, sf, out, stats);
    } else {
        if ((rc = set_smem(istft_p_kernel<NF, VEC, false>, smem)) != ADV_OK) return rc;
        istft_p_kernel<NF, VEC, false><<<grid, kThreads, smem, s>>>(p->d, tl, (int)total, X, sb, st, sf, out, stats);
    }
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

template <int NF>
static int launch_istft_pv(const adv_plan* p, const float2* X, int64_t sb, int64_t st, int64_t sf, int batch,
                           float* out, double* stats, cudaStream_t s) {
    const Tiling tl = choose_tiling(p, batch, 2, istft_balanced());
    // widest gather the geometry allows: hop, row pitch and base address must keep VEC-sample groups aligned
    const uintptr_t base = reinterpret_cast<uintptr_t>(out);
    int vec = 1;
    if (p->d.hop % 4 == 0 && p->d.n_out % 4 == 0 && base % 16 == 0) vec = 4;
    else if (p->d.hop % 2 == 0 && p->d.n_out % 2 == 0 && base % 8 == 0) vec = 2;
    if (vec == 4) return launch_istft_p<NF, 4>(p, tl, X, sb, st, sf, batch, out, stats, s);
    if (vec == 2) return launch_istft_p<NF, 2>(p, tl, X, sb, st, sf, batch, out, stats, s);
    return launch_istft_p<NF, 1>(p, tl, X, sb, st, sf, batch, out, stats, s);
}

template <bool RECT, int HS>
static int launch_istft_w512(const adv_plan* p, const float2* X, int64_t sb, int64_t st, int batch, float* out,
                             double* stats, cudaStream_t s) {
    const Tiling tl = choose_tiling(p, batch, 2, istft_balanced());  // (same tiling as istft_p_kernel: the stats layout is unchanged)
    const size_t smem = IWCfg::bytes(p->d.hop, p->d.whi - p->d.wlo);
    const long total = (long)tl.tiles * batch;
    if (total > 0x7fffffffL) return ADV_ERR_UNSUPPORTED;
    auto kernel = istft_w512_kernel<RECT, HS>;
    int rc = set_smem(kernel, smem);
    if (rc != ADV_OK) return rc;
    const int resident = resident_memo(kernel, kWideThreads, smem, 2);
    const long slots = (long)resident * sm_count();
    const int grid = (int)(total < slots ? total : slots);
    ADV_CUDA_CHECK(launch_pdl(kernel, grid, kWideThreads, smem, s, p->d, tl, (int)total, X, sb, st, out, stats));
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

int launch_istft(const adv_plan* p, const float2* X, int64_t sb, int64_t st, int64_t sf, int batch, float* out,
                 double* stats, cudaStream_t s) {
    if (istft4_slots(p, batch) > 0) return launch_istft4(p, X, sb, st, sf, batch, out, stats, s);
    if (istft5_slots(p, batch) > 0) {   // n_fft 1024 streaming kernel; the plan's statistics layout is its own
        const int rc5 = launch_istft5(p, X, sb, st, sf, batch, out, stats, s);
        if (rc5 != ADV_ERR_UNSUPPORTED) return rc5;
        if (stats != nullptr) return ADV_ERR_INVALID;   // (an output row that is not 8-byte aligned)
    }
    {
        const int rc3 = launch_istft3(p, X, sb, st, sf, batch, out, stats, s);
        if (rc3 != ADV_ERR_UNSUPPORTED) return rc3;
    }
#ifdef ADV_AB
    static const bool v2 = ADV_AB_ENV("ADV_ISTFT_V2") != nullptr;  // A/B switch: the one-tile-per-CTA kernel
    if (v2 && p->d.n_fft == 512) return launch_istft_nf<512>(p, X, sb, st, sf, batch, out, stats, s);
    if (v2 && p->d.n_fft == 1024) return launch_istft_nf<1024>(p, X, sb, st, sf, batch, out, stats, s);
    // generation 2, n_fft 512, contiguous rows, 4-sample groups aligned (reachable with ADV_GEN3=0 ADV_GEN4=0): the
    // wide-unit kernel (ADV_ISTFT_W512=0 selects the narrow-unit istft_p_kernel)
    static const char* w512_env = ADV_AB_ENV("ADV_ISTFT_W512");
    static const bool w512_on = !(w512_env && w512_env[0] == '0');
    if (w512_on && p->d.n_fft == 512 && sf == 1 && p->d.hop % 4 == 0 && p->d.n_out % 4 == 0 &&
        reinterpret_cast<uintptr_t>(out) % 16 == 0 && IWCfg::bytes(p->d.hop, p->d.whi - p->d.wlo) <= 110 * 1024) {
        // register-local overlap-add where hop is a multiple of 32 (hop 128: 42.8 -> 33.3 us against the generic strip path)
        if (p->d.rect_full && p->d.hop == 160) return launch_istft_w512<true, 5>(p, X, sb, st, batch, out, stats, s);
        if (p->d.rect_full && p->d.hop == 128) return launch_istft_w512<true, 4>(p, X, sb, st, batch, out, stats, s);
        if (p->d.rect_full && p->d.hop == 256) return launch_istft_w512<true, 8>(p, X, sb, st, batch, out, stats, s);
        return p->d.rect_full ? launch_istft_w512<true, 0>(p, X, sb, st, batch, out, stats, s)
                              : launch_istft_w512<false, 0>(p, X, sb, st, batch, out, stats, s);
    }
#endif
    // What is left for generation 2 in the product build: n_fft 512 with strided rows / odd hops / unaligned outputs
    // (istft_p_kernel) and n_fft 1024 geometries outside generation 3's domain (odd hop, oversized strips).
    // n_fft 1024 (one unit per warp, 8 units per CTA): the persistent kernel measured slower than the
    // one-tile-per-CTA kernel (51 vs 47 us on 64 x 5 s clips) - it spills around the row prefetch
    if (p->d.n_fft == 512) return launch_istft_pv<512>(p, X, sb, st, sf, batch, out, stats, s);
#ifdef ADV_AB
    static const bool p1024 = ADV_AB_ENV("ADV_ISTFT_P1024") != nullptr;
    if (p1024) return launch_istft_pv<1024>(p, X, sb, st, sf, batch, out, stats, s);
#endif
    return launch_istft_nf<1024>(p, X, sb, st, sf, batch, out, stats, s);
}

template <int NF, int MODE, bool FROM_SPEC>
static int launch_explain_inst(const adv_plan* p, const Tiling& tl, const float* wav, int64_t wav_stride,
                               const float2* X, int64_t sb, int64_t st, int64_t sf, const float* mask, int Fm,
                               int Tm, int drop, int batch, float* rel, float* irr, double* stats, cudaStream_t s) {
    const size_t smem = Cfg<NF>::explain_bytes(p->d.hop, p->d.whi - p->d.wlo, FROM_SPEC, tl.hops_per_tile * p->d.hop);
    int rc = set_smem(explain_kernel<NF, MODE, FROM_SPEC>, smem);
    if (rc != ADV_OK) return rc;
    dim3 grid(tl.tiles, batch);
    explain_kernel<NF, MODE, FROM_SPEC><<<grid, kThreads, smem, s>>>(p->d, tl, wav, wav_stride, X, sb, st, sf,
                                                                     mask, Fm, Tm, drop, rel, irr, stats);
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

template <int MODE, bool FROM_SPEC>
static int launch_explain_wide(const adv_plan* p, const Tiling& tl, const float* wav, int64_t wav_stride,
                               const float2* X, int64_t sb, int64_t st, int64_t sf, const float* mask, int Fm,
                               int Tm, int drop, int batch, float* rel, float* irr, double* stats, cudaStream_t s) {
    const size_t smem = WideCfg::bytes(p->d.hop, p->d.whi - p->d.wlo, FROM_SPEC, tl.hops_per_tile * p->d.hop);
    dim3 grid(tl.tiles, batch);
    int rc;
    if (p->d.rect_full) {
        if ((rc = set_smem(explain_w512_kernel<MODE, FROM_SPEC, true>, smem)) != ADV_OK) return rc;
        explain_w512_kernel<MODE, FROM_SPEC, true><<<grid, kWideThreads, smem, s>>>(p->d, tl, wav, wav_stride, X, sb, st,
                                                                                   sf, mask, Fm, Tm, drop, rel, irr, stats);
    } else {
        if ((rc = set_smem(explain_w512_kernel<MODE, FROM_SPEC, false>, smem)) != ADV_OK) return rc;
        explain_w512_kernel<MODE, FROM_SPEC, false><<<grid, kWideThreads, smem, s>>>(p->d, tl, wav, wav_stride, X, sb, st,
                                                                                    sf, mask, Fm, Tm, drop, rel, irr, stats);
    }
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

int launch_explain(const adv_plan* p, const float* wav, int64_t wav_stride, const float2* X, int64_t sb,
                   int64_t st, int64_t sf, const float* mask, int Fm, int Tm, int mode, int batch, float* rel,
                   float* irr, double* stats, cudaStream_t s) {
    if (explain4_slots(p, batch) > 0) {   // the plan's statistics layout is the generation-4 kernel's (adv_plan_tiles)
        if (X == nullptr) return launch_explain4(p, wav, wav_stride, mask, Fm, Tm, mode, batch, rel, irr, stats, s);
        if (stats != nullptr) return ADV_ERR_INVALID;   // spectrum input wants a plan created with n_in = 0
    }
    if (explain5_slots(p, batch) > 0) {   // n_fft 1024 streaming kernel (the plan's statistics layout is its own)
        if (X == nullptr) {
            const int rc5 = launch_explain5(p, wav, wav_stride, mask, Fm, Tm, mode, batch, rel, irr, stats, s);
            if (rc5 != ADV_ERR_UNSUPPORTED) return rc5;
        }
        if (stats != nullptr) return ADV_ERR_INVALID;   // spectrum input / unaligned rows want a plan created with n_in = 0
    }
    {
        const int rc3 = launch_explain3(p, wav, wav_stride, X, sb, st, sf, mask, Fm, Tm, mode, batch, rel, irr, stats, s);
        if (rc3 != ADV_ERR_UNSUPPORTED) return rc3;
    }
    const int drop = (mode & ADV_MASK_DROP_OUTSIDE) ? 1 : 0;
    mode &= 0xff;
    const Tiling tl = choose_tiling(p, batch, 1);
    const bool spec = (X != nullptr);
#ifdef ADV_AB
    static const bool narrow = ADV_AB_ENV("ADV_EXPLAIN_NARROW") != nullptr;  // keep the 16-lane-unit kernel reachable
    static const bool oldwide = ADV_AB_ENV("ADV_EXPLAIN_W512") != nullptr;   // A/B: the one-tile-per-CTA wide kernel
    if (p->d.n_fft == 512 && !narrow && !oldwide && !spec && !drop && p->d.hop % 4 == 0 && p->d.n_out % 4 == 0 &&
        (reinterpret_cast<uintptr_t>(rel) & 15) == 0 && (reinterpret_cast<uintptr_t>(irr) & 15) == 0 &&
        (tl.hops_per_tile * p->d.hop) % 4 == 0) {
        const size_t smem = PWideCfg::bytes(p->d.hop, p->d.whi - p->d.wlo);
        const long total = (long)tl.tiles * batch;
        if (smem <= 227 * 1024 && total <= 0x7fffffffL) {
            const int grid = (int)(total < sm_count() ? total : sm_count());
            int rc;
#define ADV_PWIDE(MODE, RECT)                                                                                   \
    do {                                                                                                        \
        if ((rc = set_smem(explain_p512_kernel<MODE, RECT>, smem)) != ADV_OK) return rc;                        \
        ADV_CUDA_CHECK(launch_pdl(explain_p512_kernel<MODE, RECT>, grid, kWideThreads, smem, s, p->d, tl,       \
                                  (int)total, wav, wav_stride, mask, Fm, Tm, rel, irr, stats));                 \
    } while (0)
            if (mode == ADV_MASK_LOG1P) { if (p->d.rect_full) ADV_PWIDE(ADV_MASK_LOG1P, true); else ADV_PWIDE(ADV_MASK_LOG1P, false); }
            else { if (p->d.rect_full) ADV_PWIDE(ADV_MASK_LINEAR, true); else ADV_PWIDE(ADV_MASK_LINEAR, false); }
#undef ADV_PWIDE
            ADV_CUDA_CHECK(cudaGetLastError());
            return ADV_OK;
        }
    }
#endif
#ifndef ADV_AB
    if (p->d.n_fft == 512) {
#else
    if (p->d.n_fft == 512 && !narrow) {
#endif
#define ADV_WIDE(MODE, SPEC) \
    return launch_explain_wide<MODE, SPEC>(p, tl, wav, wav_stride, X, sb, st, sf, mask, Fm, Tm, drop, batch, rel, irr, stats, s)
        if (mode == ADV_MASK_LOG1P) { if (spec) ADV_WIDE(ADV_MASK_LOG1P, true); else ADV_WIDE(ADV_MASK_LOG1P, false); }
        else { if (spec) ADV_WIDE(ADV_MASK_LINEAR, true); else ADV_WIDE(ADV_MASK_LINEAR, false); }
#undef ADV_WIDE
    }
#define ADV_EXPLAIN(NF, MODE, SPEC) \
    return launch_explain_inst<NF, MODE, SPEC>(p, tl, wav, wav_stride, X, sb, st, sf, mask, Fm, Tm, drop, batch, rel, irr, stats, s)
#ifdef ADV_AB
    if (p->d.n_fft == 512) {
        if (mode == ADV_MASK_LOG1P) { if (spec) ADV_EXPLAIN(512, ADV_MASK_LOG1P, true); else ADV_EXPLAIN(512, ADV_MASK_LOG1P, false); }
        else { if (spec) ADV_EXPLAIN(512, ADV_MASK_LINEAR, true); else ADV_EXPLAIN(512, ADV_MASK_LINEAR, false); }
    } else
#endif
    {
        if (mode == ADV_MASK_LOG1P) { if (spec) ADV_EXPLAIN(1024, ADV_MASK_LOG1P, true); else ADV_EXPLAIN(1024, ADV_MASK_LOG1P, false); }
        else { if (spec) ADV_EXPLAIN(1024, ADV_MASK_LINEAR, true); else ADV_EXPLAIN(1024, ADV_MASK_LINEAR, false); }
    }
#undef ADV_EXPLAIN
}

}  // namespace adv
