// Generation-3 STFT / iSTFT / fused explain kernels for sm_100a: the shuffle-free 8 x 8 x 8 warp FFT of fft3.cuh.
//
// Reference call sites (as for transform_kernels.cu): torch.stft + abs + angle (audioprocessor.py:102-110),
// torch.istft (audioprocessor.py:123-129), STFT -> mask / (1 - mask) -> 2 x iSTFT (LMAC_metrics.py:136-157 log1p,
// loss_function.py:36-47 linear).
//
// One warp = one 512-point complex transform, 16 values per lane:
//   n_fft 512 : two adjacent real frames ride in one complex transform (split / merge are register-local);
//   n_fft 1024: ONE real frame = one 512-point complex transform of z[m] = x[2m] + i x[2m+1] plus a register-local
//               twiddle pass (fft3.cuh r1024_post / r1024_pre).  The reference's default geometry (1024 / 322 / 644)
//               therefore runs on the same 64 - 96 register warps as the benchmark geometry instead of the
//               32-values-per-lane units of the first generation (128 - 250 registers).
// Kernel skeletons (warp-autonomous STFT with bulk-async slices, persistent overlap-add kernels with private strips
// and a 4-sample gather, double-buffered cp.async mask tiles) are those of the generation-2 kernels; what changed is
// the transform, the bin ownership (32 consecutive bins per load / store instruction) and that no instruction in
// the hot loops is a shuffle.  Strip origins are rounded down to a multiple of 4 samples per unit ("rem" shift), so
// the 128-bit gather also serves hops that are only even (322).
#include <type_traits>
#include "transform_common.cuh"
#include "fft3.cuh"

namespace adv {

constexpr int kT3 = 512;  // threads of the overlap-add kernels (16 warps)

template <int NF>
struct TW3 {
    static constexpr int N = NF == 512 ? f3::TW1024_OFF : f3::TW_TOTAL;  // float2 entries staged in shared memory
};
template <int NF, int NT>
__device__ __forceinline__ void stage_tw3(float2* tw_s, float* win_s, const PlanDev& P, bool want_window) {
    for (int i = threadIdx.x; i < TW3<NF>::N / 2; i += NT) cp_async16(tw_s + 2 * i, P.tw3 + 2 * i);
    if (want_window)
        for (int i = threadIdx.x; i < NF / 4; i += NT) cp_async16(win_s + 4 * i, P.window + 4 * i);
}

// lane's bin of slot i (i < 8) as base + 64 * (i & 3)
__device__ __forceinline__ int q1_lane(int l) { return l == 0 ? 32 : 64 - l; }

// ------------------------------------------------------------------------------------------------------------------
// STFT: warp-autonomous, item = two adjacent frames of one clip
// ------------------------------------------------------------------------------------------------------------------
template <int NF, bool VEC>
struct S3Cfg {
    static constexpr int WARPS = kThreads / 32;
    static __host__ __device__ size_t seg_floats(int hop) { return ((size_t)hop + NF + 8 + 3) & ~size_t(3); }
    static size_t bytes(int hop, bool rect) {
        return al16(sizeof(float2) * TW3<NF>::N) + al16(sizeof(float) * WARPS * f3::Scr<VEC>::FLOATS) +
               (rect ? 0 : al16(sizeof(float) * NF)) + al16(sizeof(float) * WARPS * seg_floats(hop)) +
               al16(sizeof(uint64_t) * WARPS);
    }
};

template <bool MAG, bool PHASE>
__device__ __forceinline__ void store_bin(float2* __restrict__ X, float* __restrict__ mag, float* __restrict__ phase,
                                          size_t idx, float2 x) {
    X[idx] = x;
    if (MAG) mag[idx] = fast_abs2(x);
    if (PHASE) phase[idx] = fast_atan2f(x.y, x.x);
}

template <int NF, bool MAG, bool PHASE, bool RECT, bool VEC, bool ZP>
#ifndef ADV_STFT3_VEC_CTAS
#define ADV_STFT3_VEC_CTAS 3
#endif
#ifndef ADV_STFT3_SHARED_ROWS
#define ADV_STFT3_SHARED_ROWS 1
#endif
#ifndef ADV_STFT3_DYN
#define ADV_STFT3_DYN 0   // 1: items drawn from per-group counters instead of the static round-robin (A/B, see the kernel)
#endif
__global__ void __launch_bounds__(kThreads, VEC ? ADV_STFT3_VEC_CTAS : 4)
stft3_kernel(PlanDev P, const float* __restrict__ wav, int64_t wav_stride, int total_items, int items_per_clip,
             float2* __restrict__ X, float* __restrict__ mag, float* __restrict__ phase, int* __restrict__ work, int groups) {
    using C = S3Cfg<NF, VEC>;
    constexpr int F = NF / 2 + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Carver cv{smem_raw};
    float2* tw_s = cv.take<float2>(TW3<NF>::N);
    float* scratch = cv.take<float>(C::WARPS * f3::Scr<VEC>::FLOATS);
    float* win_s = RECT ? nullptr : cv.take<float>(NF);
    float* seg_all = cv.take<float>(C::WARPS * C::seg_floats(P.hop));
    uint64_t* bars = cv.take<uint64_t>(C::WARPS);

    const int tid = threadIdx.x, l = tid & 31;
    const int w = __shfl_sync(0xffffffffu, tid >> 5, 0);   // (tells the compiler the warp index is warp-uniform)
    float* seg = seg_all + (size_t)w * C::seg_floats(P.hop);
    uint64_t* bar = bars + w;
    const int seglen = P.hop + NF;
#if ADV_STFT3_DYN
    // A/B variant (-DADV_STFT3_DYN=1): every warp starts on item blockIdx * WARPS + w and then DRAWS its next items from
    // a counter instead of walking the static round-robin.  Why it exists: at 1 024 clips with all three outputs the static
    // schedule leaves SMs idle for 16 % of the launch (ncu, profiles/r02s_*: per-SM active cycles 639 k ... 894 k of 907 k -
    // SMs differ in their write path to HBM).  The CTAs form `groups` equal groups (blockIdx mod groups, groups = the largest
    // divisor of the grid up to 16, members spread over the chip); group g draws from its own counter - 256 bytes apart: 12
    // counters in ONE L2 line serialise like one (17.2 -> 23.7 us per 64 clips, 218 -> 460 us per 1 024) - and owns the items
    // n_warps + g + groups * j.  Draws are issued two items ahead (lane 0, broadcast when needed).
    // Measured (profiles/r02v_stft_dyn_ab.jsonl): X + |X| + angle at 1 024 clips 455 -> 426 us (66 -> 70 % of HBM peak), no
    // change at 256 clips (111.9 / 111.5 us), SLOWER at 64 clips (28.0 -> 30.4 us; X only 17.6 -> 20.5 us: the opening draws
    // and a non-uniform trip count cost more than 3.6 items per warp can win back).  The static schedule stays the default.
    const int n_warps = gridDim.x * C::WARPS;
    const int grp = blockIdx.x % groups;
    auto draw = [&]() -> int { return l == 0 ? n_warps + grp + groups * atomicAdd(work + grp * kWorkPad, 1) : 0; };

    if (l == 0) mbar_init(bar, 1);
    stage_tw3<NF, kThreads>(tw_s, win_s, P, !RECT);
    pdl_launch_dependents();
    cp_async_wait_all();
    __syncthreads();  // tables staged, barriers initialised
    pdl_wait();

    float* my = scratch + w * f3::Scr<VEC>::FLOATS;
    const int q1 = q1_lane(l);
    int item = blockIdx.x * C::WARPS + w;
    int d1 = 0, d2 = 0;
    bool have = item < total_items;
    int b = 0, t0 = 0, shift = 0;
    if (have) {
        d1 = draw();
        d2 = draw();
        b = item / items_per_clip;
        t0 = (item - b * items_per_clip) * 2;
        shift = stage_segment_async<32>(seg, seglen, wav + (size_t)b * wav_stride, t0 * P.hop - NF / 2, P.n_in, bar, l, ZP);
    }
    constexpr bool active = true;

    for (uint32_t it = 0; have; ++it) {
        __syncwarp();  // plain-load part of the slice visible to the warp
        mbar_wait(bar, it & 1);
        const int cur_b = b, fa = t0, cur_shift = shift;
        const int next = __shfl_sync(0xffffffffu, d1, 0);
        const bool more = next < total_items;   // (warp-uniform)
        have = more;
        d1 = d2;
        d2 = more ? draw() : 0;
        auto request_next = [&]() {   // (called from inside the forward transform: every lane has consumed its samples)
            if (more) {
                item = next;
                b = item / items_per_clip;
                t0 = (item - b * items_per_clip) * 2;
                shift = stage_segment_async<32>(seg, seglen, wav + (size_t)b * wav_stride, t0 * P.hop - NF / 2, P.n_in, bar,
                                                l, ZP);
            }
        };
#else
    const int stride = gridDim.x * C::WARPS;
    // per-warp trip count: a warp leaves the loop after its last item (and requests no slice it will not consume).  Until
    // r02zz the count was CTA-uniform and a warp past the list repeated the last item with its stores switched off: at 64
    // clips that is 4 rounds billed for 3.6.
    const int first = blockIdx.x * C::WARPS + w;
    const int n_iter = first < total_items ? (total_items - first + stride - 1) / stride : 0;

    if (l == 0) mbar_init(bar, 1);
    stage_tw3<NF, kThreads>(tw_s, win_s, P, !RECT);
    pdl_launch_dependents();
    cp_async_wait_all();
    __syncthreads();  // tables staged, barriers initialised
    pdl_wait();

    float* my = scratch + w * f3::Scr<VEC>::FLOATS;
    const int q1 = q1_lane(l);
    int item = min(first, total_items - 1);
    int b = item / items_per_clip, t0 = (item - b * items_per_clip) * 2;
    int shift = 0;
    if (n_iter > 0)
        shift = stage_segment_async<32>(seg, seglen, wav + (size_t)b * wav_stride, t0 * P.hop - NF / 2, P.n_in, bar, l, ZP);

    for (int it = 0; it < n_iter; ++it) {
        __syncwarp();  // plain-load part of the slice visible to the warp
        mbar_wait(bar, it & 1);
        const int cur_b = b, fa = t0, cur_shift = shift;
        constexpr bool active = true;
        const bool more = it + 1 < n_iter;  // (warp-uniform)
        auto request_next = [&]() {   // (called from inside the forward transform: every lane has consumed its samples)
            if (more) {
                item = first + (it + 1) * stride;
                b = item / items_per_clip;
                t0 = (item - b * items_per_clip) * 2;
                shift = stage_segment_async<32>(seg, seglen, wav + (size_t)b * wav_stride, t0 * P.hop - NF / 2, P.n_in, bar,
                                                l, ZP);
            }
        };
#endif
        if constexpr (NF == 512) {
            float2 v[16];
            {
                const float* sa = seg + cur_shift + l;
#if ADV_STFT3_SHARED_ROWS
                // hop = 32 HS: frame b's row j is frame a's row j + HS - the two frames need 16 + HS distinct rows of 32
                // samples, not 32 (shared-memory instructions are this kernel's top stall: mio_throttle)
                auto load_rows = [&](auto hs) {
                    constexpr int HS = decltype(hs)::value;
                    float r[16 + HS];
#pragma unroll
                    for (int j = 0; j < 16 + HS; ++j) r[j] = sa[32 * j];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float ww = RECT ? 1.0f : win_s[32 * j + l];
                        v[j] = make_float2(r[j] * ww, r[j + HS] * ww);
                    }
                };
                if (P.hop == 160) load_rows(std::integral_constant<int, 5>{});
                else if (P.hop == 128) load_rows(std::integral_constant<int, 4>{});
                else if (P.hop == 256) load_rows(std::integral_constant<int, 8>{});
                else
#endif
                {
                    const float* sb = sa + P.hop;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float ww = RECT ? 1.0f : win_s[32 * j + l];
                        v[j] = make_float2(sa[32 * j] * ww, sb[32 * j] * ww);
                    }
                }
            }
            f3::fft_forward<VEC>(v, l, tw_s, my, request_next);
            float2 xa[9], xb[9];
            f3::split(v, l, xa, xb);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int t = fa + half;
                if (!active || t >= P.T) continue;
                const float2* x = half ? xb : xa;
                const size_t row = ((size_t)cur_b * P.T + t) * F;
#pragma unroll
                for (int i = 0; i < 4; ++i) store_bin<MAG, PHASE>(X, mag, phase, row + l + 64 * i, x[i]);
#pragma unroll
                for (int i = 0; i < 4; ++i) store_bin<MAG, PHASE>(X, mag, phase, row + q1 + 64 * i, x[4 + i]);
                if (l == 0) store_bin<MAG, PHASE>(X, mag, phase, row + 256, x[8]);
            }
        } else {
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                float2 v[16];
                {
                    const float* s = seg + cur_shift + half * P.hop + 2 * l;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float2 x = *reinterpret_cast<const float2*>(s + 64 * j);
                        if (RECT) {
                            v[j] = x;
                        } else {
                            const float2 ww = *reinterpret_cast<const float2*>(win_s + 64 * j + 2 * l);
                            v[j] = make_float2(x.x * ww.x, x.y * ww.y);
                        }
                    }
                }
                const int t = fa + half;
                f3::fft_forward<VEC>(v, l, tw_s, my, [&] { if (half == 1) request_next(); });
                float2 xk[9], xm[9];
                f3::r1024_post(v, l, tw_s, xk, xm);
                if (active && t < P.T) {
                    const size_t row = ((size_t)cur_b * P.T + t) * F;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        store_bin<MAG, PHASE>(X, mag, phase, row + l + 64 * i, xk[i]);
                        store_bin<MAG, PHASE>(X, mag, phase, row + 512 - l - 64 * i, xm[i]);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        store_bin<MAG, PHASE>(X, mag, phase, row + q1 + 64 * i, xk[4 + i]);
                        store_bin<MAG, PHASE>(X, mag, phase, row + 512 - q1 - 64 * i, xm[4 + i]);
                    }
                    if (l == 0) store_bin<MAG, PHASE>(X, mag, phase, row + 256, xk[8]);
                }
            }
        }
    }
#if ADV_STFT3_DYN
    // Retire: the warp's outstanding draws must have returned before it reports (a draw still in flight would land on
    // the counter AFTER the reset below); the last warp of the launch leaves both counters at zero for the next one.
    if (l == 0) {
        asm volatile("" ::"r"(d1), "r"(d2) : "memory");
        __threadfence();
        if (atomicAdd(work + kWorkGroups * kWorkPad, 1) == n_warps - 1) {
#pragma unroll
            for (int i = 0; i <= kWorkGroups; ++i) work[i * kWorkPad] = 0;
            __threadfence();
        }
    }
#endif
}

// ------------------------------------------------------------------------------------------------------------------
// overlap-add geometry shared by the iSTFT and the fused kernel
//   unit u of a tile owns FPU consecutive frames starting at fa = t_lo + FPU * u and a private strip whose origin (in
//   padded sample coordinates) is  o_u = (fa * hop + wlo4) & ~3,  wlo4 = wlo & ~3;  rem_u = (fa * hop + wlo4) & 3.
//   Frame sample n (>= wlo4) of the unit's j-th frame sits at strip position rem_u + j * hop + n - wlo4.
// ------------------------------------------------------------------------------------------------------------------
struct OlaGeo {
    int hop, wlo4, sup;   // sup = whi - wlo4 (window support measured from wlo4)
    int lb;               // strip length the gather may read (multiple of 4, >= rem + (FPU - 1) * hop + sup)
    int pitch;            // strip pitch in elements
};
template <int FPU>
__host__ __device__ inline OlaGeo ola_geo(int hop, int wlo, int whi) {
    OlaGeo g;
    g.hop = hop;
    g.wlo4 = wlo & ~3;
    g.sup = whi - g.wlo4;
    g.lb = ((FPU - 1) * hop + g.sup + 3 + 3) & ~3;
    g.pitch = g.lb + 4;
    return g;
}

// number of the highest unit whose strip starts at or before padded position pp (may be negative: none)
template <int UNITS>
__device__ __forceinline__ int top_unit(int pp, int base0, int ustep, float inv_ustep) {
    const int num = (pp | 3) - base0;
    const int q = (int)(((float)num + 0.5f) * inv_ustep);  // exact for 0 <= num < 2^20
    return num < 0 ? -1 : min(UNITS - 1, q);
}

// ------------------------------------------------------------------------------------------------------------------
// iSTFT: persistent, 16 units x 2 frames per tile
// ------------------------------------------------------------------------------------------------------------------
template <int NF>
struct I3Cfg {
    static constexpr int UNITS = 16;
    static size_t bytes(int hop, int wlo, int whi, bool rect) {
        const OlaGeo g = ola_geo<2>(hop, wlo, whi);
        return al16(sizeof(float2) * TW3<NF>::N) + al16(sizeof(float) * UNITS * f3::Scr<false>::FLOATS) +
               (rect ? 0 : al16(sizeof(float) * NF)) + al16(sizeof(float) * UNITS * g.pitch) +
               al16(sizeof(double) * 2 * (kT3 / 32));
    }
};

// HS > 0 (n_fft 512 only): rectangular full window with hop = 32 * HS - frame b's sample 32 j + l lands on strip position
// 32 (j + HS) + l, the SAME lane as frame a's sample: the two frames are overlap-added in registers.
template <int NF, bool RECT, int HS, bool CONTIG, int GV>
__global__ void __launch_bounds__(kT3, 2)
istft3_kernel(PlanDev P, Tiling TL, int total_tiles, const float2* __restrict__ X, int64_t sb, int64_t st, int64_t sf,
              float* __restrict__ out, double* __restrict__ stats) {
    using C = I3Cfg<NF>;
    constexpr int UNITS = C::UNITS, NT = kT3;
    constexpr bool FIX = NF == 512 && RECT && HS > 0;
    static_assert(!(HS > 0) || (NF == 512 && RECT), "register-local overlap-add: n_fft 512, rectangular window");
    const OlaGeo G = FIX ? ola_geo<2>(32 * HS, 0, NF) : ola_geo<2>(P.hop, P.wlo, P.whi);
    const int hop = G.hop;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Carver cv{smem_raw};
    float2* tw_s = cv.take<float2>(TW3<NF>::N);
    float* scratch = cv.take<float>(UNITS * f3::Scr<false>::FLOATS);
    float* win_s = RECT ? nullptr : cv.take<float>(NF);
    float* pb = cv.take<float>(UNITS * G.pitch);
    double* red = cv.take<double>(2 * (NT / 32));

    const int tid = threadIdx.x, u = tid >> 5, l = tid & 31;
    stage_tw3<NF, NT>(tw_s, win_s, P, !RECT);
    pdl_launch_dependents();

    const int q1 = q1_lane(l);
    const int64_t sfe = CONTIG ? 1 : sf;
    // next tile's rows -> L2 (no registers held across the gather): lane l touches line l of each row
    auto prefetch_rows = [&](int t_id) {
        const int bb = t_id / TL.tiles;
        const TileGeom gg = tile_geom<NF>(P, TL, t_id - bb * TL.tiles);
        const int fa = gg.t_lo + 2 * u;
        if (CONTIG && fa + 1 <= gg.t_hi && l * 16 < NF / 2 + 1) {
            const float2* xa_p = X + (size_t)bb * sb + (size_t)fa * st + l * 16;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(xa_p));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(xa_p + st));
        }
    };
    int tile = blockIdx.x;
    cp_async_wait_all();
    __syncthreads();
    pdl_wait();  // (everything above touches plan constants only)

    float* my = scratch + u * f3::Scr<false>::FLOATS;
    float* pbu = pb + u * G.pitch;
    const int ustep = 2 * hop;
    const float inv_ustep = 1.0f / (float)ustep;
    const float2 zero2 = make_float2(0.f, 0.f);

    for (;;) {
        const int b = tile / TL.tiles, cur_tile = tile - b * TL.tiles;
        const TileGeom g = tile_geom<NF>(P, TL, cur_tile);
        const int next = tile + gridDim.x;
        const int fa = g.t_lo + 2 * u;
        const int base0 = g.t_lo * hop + G.wlo4;               // unrounded origin of unit 0
        const int rem = FIX ? 0 : ((fa * hop + G.wlo4) & 3);
        {   // No branch on the unit's frame range: frames past the tile's last one are transformed as zero spectra
            // (a few idle units in the last tile of a clip), which keeps the transforms and their __syncwarp()s in
            // provably convergent code - a warp-uniform but thread-index-derived condition costs a WARPSYNC /
            // ENDCOLLECTIVE / NOP sequence around every warp-level sync it encloses.
            const bool va = fa <= g.t_hi, vb = fa + 1 <= g.t_hi;
            const float2* xa_p = X + (size_t)b * sb + (size_t)fa * st;
            if constexpr (NF == 512) {
                float2 xa[9], xb[9];
                const float2* xb_p = xa_p + st;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    xa[i] = va ? __ldg(xa_p + (l + 64 * i) * sfe) : zero2;
                    xa[4 + i] = va ? __ldg(xa_p + (q1 + 64 * i) * sfe) : zero2;
                    xb[i] = vb ? __ldg(xb_p + (l + 64 * i) * sfe) : zero2;
                    xb[4 + i] = vb ? __ldg(xb_p + (q1 + 64 * i) * sfe) : zero2;
                }
                xa[8] = (l == 0 && va) ? __ldg(xa_p + 256 * sfe) : zero2;
                xb[8] = (l == 0 && vb) ? __ldg(xb_p + 256 * sfe) : zero2;
                float2 v[16];
                f3::merge(v, l, xa, xb);
                f3::fft_inverse<false>(v, l, tw_s, my);
                // private strip of the unit: frame a at [rem, rem + sup), frame b hop further
                if constexpr (FIX) {
#pragma unroll
                    for (int j = 0; j < 16 + HS; ++j) {
                        float o = j < 16 ? v[j < 16 ? j : 0].x : 0.0f;
                        if (j >= HS) o += v[j >= HS ? j - HS : 0].y;
                        pbu[j * 32 + l] = o;
                    }
                    if (l < G.lb - (16 + HS) * 32) pbu[(16 + HS) * 32 + l] = 0.0f;  // tail the 4-sample gather may touch
                } else {
                    const int c0 = l - G.wlo4, ovl = G.sup - hop;
                    float* pa = pbu + rem;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int k = 32 * j + c0;
                        if ((unsigned)k < (unsigned)G.sup) pa[k] = RECT ? v[j].x : v[j].x * win_s[32 * j + l];
                    }
                    if (ovl < 0)  // degenerate (hop > support, only legal for single-frame plans): clear the gap
                        for (int k = G.sup + l; k < hop; k += 32) pa[k] = 0.0f;
                    if (l < 8) {  // the <= 3 elements before frame a and the tail the gather may touch
                        if (l < rem) pbu[l] = 0.0f;
                        const int k = rem + hop + G.sup + l;
                        if (k < G.lb) pbu[k] = 0.0f;
                    }
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int k = 32 * j + c0;
                        if ((unsigned)k < (unsigned)G.sup) {
                            const float old = k < ovl ? pa[hop + k] : 0.0f;
                            pa[hop + k] = RECT ? v[j].y + old : fmaf(v[j].y, win_s[32 * j + l], old);
                        }
                    }
                }
            } else {
                const int ovl = G.sup - hop;
                float* pa = pbu + rem;
                if (l < 8) {
                    if (l < rem) pbu[l] = 0.0f;
                    const int k = rem + hop + G.sup + l;
                    if (k < G.lb) pbu[k] = 0.0f;
                }
                if (ovl < 0)
                    for (int k = G.sup + l; k < hop; k += 32) pa[k] = 0.0f;
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    const bool live = half == 0 ? va : vb;
                    const float2* xp = xa_p + half * st;
                    float2 xk[9], xm[9];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        xk[i] = live ? __ldg(xp + (l + 64 * i) * sfe) : zero2;
                        xm[i] = live ? __ldg(xp + (512 - l - 64 * i) * sfe) : zero2;
                        xk[4 + i] = live ? __ldg(xp + (q1 + 64 * i) * sfe) : zero2;
                        xm[4 + i] = live ? __ldg(xp + (512 - q1 - 64 * i) * sfe) : zero2;
                    }
                    xk[8] = (live && l == 0) ? __ldg(xp + 256 * sfe) : zero2;
                    xm[8] = xk[8];
                    float2 v[16];
                    f3::r1024_pre(v, l, tw_s, xk, xm);
                    f3::fft_inverse<false>(v, l, tw_s, my);
                    float* dst = pa + half * hop;
                    const int keep = half ? ovl : 0;  // frame b adds onto frame a over the first `ovl` strip samples
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int k = 64 * j + 2 * l - G.wlo4;
                        if ((unsigned)k < (unsigned)G.sup) {
                            float2 o = v[j];
                            if (!RECT) {
                                const float2 ww = *reinterpret_cast<const float2*>(win_s + 64 * j + 2 * l);
                                o.x *= ww.x;
                                o.y *= ww.y;
                            }
                            float2* d2 = reinterpret_cast<float2*>(dst + k);
                            if (k < keep) {
                                const float2 prev = *d2;
                                o.x += prev.x;
                                if (k + 1 < keep) o.y += prev.y;
                            }
                            *d2 = o;
                        }
                    }
                    __syncwarp();  // frame a's strip stores are visible before frame b's read-modify-write
                }
            }
        }
        if (next < total_tiles) prefetch_rows(next);
        __syncthreads();

        // gather: GV samples per step; strips covering padded position pp are uu = top_unit, uu - 1, ... while k < lb
        const int S = g.s1 - g.s0;
        float* orow = out + (size_t)b * P.n_out + g.s0;
        const float* erow = P.inv_env + g.s0;
        double acc0 = 0.0, acc1 = 0.0;
        if constexpr (GV == 4) {
            for (int q = tid * 4; q < S; q += NT * 4) {  // (n_out % 4 == 0 and tile length % 4 == 0: S % 4 == 0)
                const float4 e = __ldg(reinterpret_cast<const float4*>(erow + q));
                const int pp = g.p0 + q;
                int uu = top_unit<UNITS>(pp, base0, ustep, inv_ustep);
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                while (uu >= 0) {
                    const int k = pp - ((base0 + uu * ustep) & ~3);
                    if (k >= G.lb) break;
                    vadd(a, *reinterpret_cast<const float4*>(pb + uu * G.pitch + k));
                    --uu;
                }
                float sq;
                const float sm = vmul_stats(a, e, sq);
                *reinterpret_cast<float4*>(orow + q) = a;
                acc0 += (double)sm;
                acc1 += (double)sq;
            }
        } else {
            for (int q = tid; q < S; q += NT) {
                const float e = __ldg(erow + q);
                const int pp = g.p0 + q;
                int uu = top_unit<UNITS>(pp, base0, ustep, inv_ustep);
                float a = 0.0f;
                while (uu >= 0) {
                    const int k = pp - ((base0 + uu * ustep) & ~3);
                    if (k >= G.lb) break;
                    a += pb[uu * G.pitch + k];
                    --uu;
                }
                a *= e;
                orow[q] = a;
                acc0 += (double)a;
                acc1 += (double)(a * a);
            }
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

// ------------------------------------------------------------------------------------------------------------------
// fused explain: persistent, double-buffered mask tile + bulk-async waveform segment
//   n_fft 512 : 32 frames per tile, unit = 2 frames (one forward transform, two inverse transforms carrying the
//               masked-in / masked-out spectra of one frame as real / imaginary part), strips float2 (rel, irr)
//   n_fft 1024: 16 frames per tile, unit = 1 frame (one forward, two inverse transforms), two planar float strips
// ------------------------------------------------------------------------------------------------------------------
template <int NF>
struct E3Cfg {
    static constexpr int UNITS = 16, FPU = NF == 512 ? 2 : 1, FT = UNITS * FPU, MP = FT + 1, F = NF / 2 + 1;
    static constexpr int MASK_TILE = (F * MP + 3) & ~3;
    static __host__ __device__ size_t seg_floats(int hop) { return ((size_t)(FT - 1) * hop + NF + 8 + 3) & ~size_t(3); }
    static size_t bytes(int hop, int wlo, int whi, bool rect, bool from_spec) {
        const OlaGeo g = ola_geo<FPU>(hop, wlo, whi);
        return al16(sizeof(float2) * TW3<NF>::N) + al16(sizeof(float) * UNITS * f3::Scr<false>::FLOATS) +
               (rect ? 0 : al16(sizeof(float) * NF)) + al16(sizeof(float) * 2 * UNITS * g.pitch) +
               (from_spec ? 0 : al16(sizeof(float) * seg_floats(hop))) + al16(sizeof(double) * 4 * (kT3 / 32)) +
               2 * al16(sizeof(float) * MASK_TILE) + 16;
    }
};

#ifndef ADV_EXPLAIN3_MAXREG
#define ADV_EXPLAIN3_MAXREG 96   // see explain_p512_kernel: leaves registers for two normaliser CTAs next to this kernel
#endif

template <int NF, int MODE, bool RECT, bool FROM_SPEC, int GV>
__global__ void __maxnreg__(NF == 512 ? ADV_EXPLAIN3_MAXREG : 128)
explain3_kernel(PlanDev P, Tiling TL, int total_tiles, const float* __restrict__ wav, int64_t wav_stride,
                const float2* __restrict__ X, int64_t sb, int64_t st, int64_t sf, const float* __restrict__ mask, int Fm,
                int Tm, int drop, float* __restrict__ rel, float* __restrict__ irr, double* __restrict__ stats) {
    using C = E3Cfg<NF>;
    constexpr int UNITS = C::UNITS, FPU = C::FPU, FT = C::FT, MP = C::MP, F = C::F, NT = kT3;
    static_assert(!(NF == 512 && FROM_SPEC), "n_fft 512: spectrum input stays on explain_w512_kernel");
    const OlaGeo G = ola_geo<FPU>(P.hop, P.wlo, P.whi);
    const int hop = G.hop;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Carver cv{smem_raw};
    float2* tw_s = cv.take<float2>(TW3<NF>::N);
    float* scratch = cv.take<float>(UNITS * f3::Scr<false>::FLOATS);
    float* win_s = RECT ? nullptr : cv.take<float>(NF);
    float* pb = cv.take<float>(2 * UNITS * G.pitch);   // 512: float2 strips; 1024: rel strips, then irr strips
    float* seg = FROM_SPEC ? nullptr : cv.take<float>(C::seg_floats(hop));
    double* red = cv.take<double>(4 * (NT / 32));
    float* mask_all = cv.take<float>(2 * C::MASK_TILE);
    uint64_t* bar = cv.take<uint64_t>(1);

    const int tid = threadIdx.x, u = tid >> 5, l = tid & 31;
    const int seglen = (FT - 1) * hop + NF;
    int tile = blockIdx.x;
    int b = tile / TL.tiles;
    TileGeom g = tile_geom<NF>(P, TL, tile - b * TL.tiles);
    // out-of-mask bins: mask value 0 (zero extension); drop != 0 removes them from BOTH outputs (the reference crops
    // magnitude and phase to the mask's extent, LMAC_metrics.py:136-139 / loss_function.py:36-41)
    const int f_lim = drop ? Fm : 0x7fffffff, t_lim = drop ? Tm : 0x7fffffff;

    // every load of a tile that can be requested ahead of time: segment (bulk + plain edges), mask tile
    auto request_tile = [&](int bb, const TileGeom& gg, float* mask_dst) {
        if constexpr (!FROM_SPEC)
            stage_segment_async<NT>(seg, seglen, wav + (size_t)bb * wav_stride, gg.t_lo * hop - NF / 2, P.n_in, bar);
        // thread -> fixed column c and rows f0 + kRows j: source and destination advance by a constant per step
        constexpr int CW = FT, kRows = NT / CW, kTrips = (F + kRows - 1) / kRows;
        const int c = tid % CW, f0 = tid / CW;
        const int t = gg.t_lo + c;
        const bool col_ok = t < Tm && t <= gg.t_hi;
        const float* mrow = mask + (size_t)bb * Fm * Tm;
        const float* src = mrow + (size_t)f0 * Tm + (col_ok ? t : 0);
        uint32_t dst = smem_u32(mask_dst + f0 * MP + c);
#pragma unroll
        for (int j = 0; j < kTrips; ++j) {
            const int f = f0 + j * kRows;
            if (f < F) {
                const bool ok = col_ok && f < Fm;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(ok ? src : mrow), "r"(ok ? 4 : 0)
                             : "memory");
            }
            src += (size_t)kRows * Tm;
            dst += kRows * MP * 4;
        }
        cp_async_commit();
    };

    if (tid == 0) mbar_init(bar, 1);
    stage_tw3<NF, NT>(tw_s, win_s, P, !RECT);
    cp_async_commit();
    pdl_launch_dependents();
    cp_async_wait_all();
    __syncthreads();
    pdl_wait();  // (everything above touches plan constants only)
    request_tile(b, g, mask_all);
    int shift = (g.t_lo * hop - NF / 2) & 3;

    float* my = scratch + u * f3::Scr<false>::FLOATS;
    const int q1 = q1_lane(l);
    const int ustep = FPU * hop;
    const float inv_ustep = 1.0f / (float)ustep;
    const int sfe = FROM_SPEC ? (int)sf : 1;

    for (uint32_t it = 0;; ++it) {
        const float* mask_s = mask_all + (it & 1) * C::MASK_TILE;
        cp_async_wait_all();
        __syncthreads();  // mask tile + plain part of the segment visible; the previous epilogue is over
        if constexpr (!FROM_SPEC) mbar_wait(bar, it & 1);
        const int cur_b = b, cur_tile = tile - b * TL.tiles;
        const TileGeom cg = g;
        const int fa = cg.t_lo + FPU * u;
        float2 v[16];
        if constexpr (!FROM_SPEC) {
            if constexpr (NF == 512) {
                const float* sa = seg + shift + (2 * u) * hop + l;
                const float* sbp = sa + hop;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float w = RECT ? 1.0f : win_s[32 * j + l];
                    v[j] = make_float2(sa[32 * j] * w, sbp[32 * j] * w);
                }
            } else {
                const float* s = seg + shift + u * hop + 2 * l;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float2 x = *reinterpret_cast<const float2*>(s + 64 * j);
                    if (RECT) {
                        v[j] = x;
                    } else {
                        const float2 ww = *reinterpret_cast<const float2*>(win_s + 64 * j + 2 * l);
                        v[j] = make_float2(x.x * ww.x, x.y * ww.y);
                    }
                }
            }
        }
        const int next = tile + gridDim.x;
        if constexpr (!FROM_SPEC) __threadfence_block();  // the sample loads above have returned (a barrier alone does not wait for them)
        __syncthreads();  // every thread holds its samples: the segment buffer is free
        if (next < total_tiles) {
            b = next / TL.tiles;
            g = tile_geom<NF>(P, TL, next - b * TL.tiles);
            request_tile(b, g, mask_all + ((it + 1) & 1) * C::MASK_TILE);
            shift = (g.t_lo * hop - NF / 2) & 3;
        }

        const int base0 = cg.t_lo * hop + G.wlo4;
        const int rem = (fa * hop + G.wlo4) & 3;
        {   // no branch on the unit's frame range (see istft3_kernel): frames past the tile's last one carry zero gains
            if constexpr (NF == 512) {
                float2* pbu = reinterpret_cast<float2*>(pb) + u * G.pitch;
                float2 xa[9], xb[9];
                f3::fft_forward<false>(v, l, tw_s, my);
                f3::split(v, l, xa, xb);
                const int ovl = G.sup - hop;
                float2* pa = pbu + rem;
                if (ovl < 0)
                    for (int k = G.sup + l; k < hop; k += 32) pa[k] = make_float2(0.f, 0.f);
                if (l < 8) {
                    if (l < rem) pbu[l] = make_float2(0.f, 0.f);
                    const int k = rem + hop + G.sup + l;
                    if (k < G.lb) pbu[k] = make_float2(0.f, 0.f);
                }
                // fully unrolled: a rolled loop has to rotate frame b's spectrum into frame a's registers
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int t = fa + half;
                    const bool valid = t <= cg.t_hi;
                    const bool tdrop = t >= t_lim;
                    const int col = 2 * u + half;
                    const float2* x = half ? xb : xa;
                    {
                        float2 yr[9], yi[9];
#pragma unroll
                        for (int i = 0; i < 9; ++i) {
                            const int bin = i < 4 ? l + 64 * i : (i < 8 ? q1 + 64 * (i - 4) : 256);
                            const bool live = (i < 8 || l == 0) && valid;
                            const float m = live ? mask_s[bin * MP + col] : 0.0f;
                            float gr, gi;
                            mask_gains<MODE>(x[i], m, gr, gi);
                            const bool keep_i = live && !(tdrop || bin >= f_lim);
                            yr[i] = live ? make_float2(x[i].x * gr, x[i].y * gr) : make_float2(0.f, 0.f);
                            yi[i] = keep_i ? make_float2(x[i].x * gi, x[i].y * gi) : make_float2(0.f, 0.f);
                        }
                        f3::merge(v, l, yr, yi);
                    }
                    f3::fft_inverse<false>(v, l, tw_s, my);
                    float2* dst = pa + half * hop;
                    const int keep = half ? ovl : 0;
                    const int c0 = l - G.wlo4;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int k = 32 * j + c0;
                        if (RECT || (unsigned)k < (unsigned)G.sup) {
                            float2 o = v[j];
                            if (!RECT) {
                                const float w = win_s[32 * j + l];
                                o.x *= w;
                                o.y *= w;
                            }
                            if (k < keep) {
                                const float2 prev = dst[k];
                                o.x += prev.x;
                                o.y += prev.y;
                            }
                            dst[k] = o;
                        }
                    }
                    __syncwarp();  // frame a's strip stores are visible to the unit before frame b's read-modify-write
                }
            } else {
                float* pr = pb + u * G.pitch;
                float* pi_ = pb + (UNITS + u) * G.pitch;
                float2 xk[9], xm[9];
                if constexpr (FROM_SPEC) {
                    const float2* xp = X + (size_t)cur_b * sb + (size_t)fa * st;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        xk[i] = __ldg(xp + (size_t)(l + 64 * i) * sfe);
                        xm[i] = __ldg(xp + (size_t)(512 - l - 64 * i) * sfe);
                        xk[4 + i] = __ldg(xp + (size_t)(q1 + 64 * i) * sfe);
                        xm[4 + i] = __ldg(xp + (size_t)(512 - q1 - 64 * i) * sfe);
                    }
                    xk[8] = l == 0 ? __ldg(xp + (size_t)256 * sfe) : make_float2(0.f, 0.f);
                    xm[8] = xk[8];
                } else {
                    f3::fft_forward<false>(v, l, tw_s, my);
                    f3::r1024_post(v, l, tw_s, xk, xm);
                }
                const bool tdrop = fa >= t_lim, valid = fa <= cg.t_hi;
                // masked-in and masked-out spectra of the frame (bins k and 512 - k per slot)
                float2 yk[9], ym[9], zk[9], zm[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) {
                    const int k = i < 4 ? l + 64 * i : (i < 8 ? q1 + 64 * (i - 4) : 256);
                    const bool live = (i < 8 || l == 0) && valid;
                    {
                        const float m = live ? mask_s[k * MP + u] : 0.0f;
                        float gr, gi;
                        mask_gains<MODE>(xk[i], m, gr, gi);
                        const bool keep_i = live && !(tdrop || k >= f_lim);
                        yk[i] = live ? make_float2(xk[i].x * gr, xk[i].y * gr) : make_float2(0.f, 0.f);
                        zk[i] = keep_i ? make_float2(xk[i].x * gi, xk[i].y * gi) : make_float2(0.f, 0.f);
                    }
                    if (i < 8) {
                        const int km = 512 - k;
                        const float m = valid ? mask_s[km * MP + u] : 0.0f;
                        float gr, gi;
                        mask_gains<MODE>(xm[i], m, gr, gi);
                        const bool keep_i = valid && !(tdrop || km >= f_lim);
                        ym[i] = valid ? make_float2(xm[i].x * gr, xm[i].y * gr) : make_float2(0.f, 0.f);
                        zm[i] = keep_i ? make_float2(xm[i].x * gi, xm[i].y * gi) : make_float2(0.f, 0.f);
                    } else {
                        ym[i] = yk[i];
                        zm[i] = zk[i];
                    }
                }
                if (l < 8) {
                    if (l < rem) { pr[l] = 0.0f; pi_[l] = 0.0f; }
                    const int k = rem + G.sup + l;
                    if (k < G.lb) { pr[k] = 0.0f; pi_[k] = 0.0f; }
                }
#pragma unroll 1
                for (int sig = 0; sig < 2; ++sig) {
                    f3::r1024_pre(v, l, tw_s, sig ? zk : yk, sig ? zm : ym);
                    f3::fft_inverse<false>(v, l, tw_s, my);
                    float* dst = (sig ? pi_ : pr) + rem;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int k = 64 * j + 2 * l - G.wlo4;
                        if ((unsigned)k < (unsigned)G.sup) {
                            float2 o = v[j];
                            if (!RECT) {
                                const float2 ww = *reinterpret_cast<const float2*>(win_s + 64 * j + 2 * l);
                                o.x *= ww.x;
                                o.y *= ww.y;
                            }
                            *reinterpret_cast<float2*>(dst + k) = o;
                        }
                    }
                }
            }
        }

        // reciprocal envelope of this thread's first groups: requested before the barrier
        const int S = cg.s1 - cg.s0;
        const float* erow = P.inv_env + cg.s0;
        constexpr int kEnvRegs = GV == 4 ? 3 : 1;
        float4 e4[kEnvRegs];
        if constexpr (GV == 4) {
#pragma unroll
            for (int j = 0; j < kEnvRegs; ++j) {
                const int q = (tid + j * NT) * 4;
                e4[j] = q + 4 <= S ? __ldg(reinterpret_cast<const float4*>(erow + q)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        __syncthreads();

        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        float* rrow = rel + (size_t)cur_b * P.n_out + cg.s0;
        float* irow = irr + (size_t)cur_b * P.n_out + cg.s0;
        if constexpr (GV == 4) {
            int trip = 0;
            for (int q = tid * 4; q < S; q += NT * 4, ++trip) {
                float4 e;
                if (trip == 0) e = e4[0];
                else if (trip == 1) e = e4[1 % kEnvRegs];
                else if (trip == 2) e = e4[2 % kEnvRegs];
                else e = __ldg(reinterpret_cast<const float4*>(erow + q));   // (tiles longer than 6144 samples)
                const int pp = cg.p0 + q;
                int uu = top_unit<UNITS>(pp, base0, ustep, inv_ustep);
                float4 r = make_float4(0.f, 0.f, 0.f, 0.f), ir = r;
                while (uu >= 0) {
                    const int k = pp - ((base0 + uu * ustep) & ~3);
                    if (k >= G.lb) break;
                    if constexpr (NF == 512) {
                        const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float2*>(pb) + uu * G.pitch + k);
                        const float4 a = p[0], c = p[1];
                        r.x += a.x; ir.x += a.y; r.y += a.z; ir.y += a.w;
                        r.z += c.x; ir.z += c.y; r.w += c.z; ir.w += c.w;
                    } else {
                        vadd(r, *reinterpret_cast<const float4*>(pb + uu * G.pitch + k));
                        vadd(ir, *reinterpret_cast<const float4*>(pb + (UNITS + uu) * G.pitch + k));
                    }
                    --uu;
                }
                r.x *= e.x; r.y *= e.y; r.z *= e.z; r.w *= e.w;
                ir.x *= e.x; ir.y *= e.y; ir.z *= e.z; ir.w *= e.w;
                *reinterpret_cast<float4*>(rrow + q) = r;
                *reinterpret_cast<float4*>(irow + q) = ir;
                acc[0] += (double)((r.x + r.y) + (r.z + r.w));
                acc[1] += (double)(fmaf(r.x, r.x, r.y * r.y) + fmaf(r.z, r.z, r.w * r.w));
                acc[2] += (double)((ir.x + ir.y) + (ir.z + ir.w));
                acc[3] += (double)(fmaf(ir.x, ir.x, ir.y * ir.y) + fmaf(ir.z, ir.z, ir.w * ir.w));
            }
        } else {
            for (int q = tid; q < S; q += NT) {
                const float e = __ldg(erow + q);
                const int pp = cg.p0 + q;
                int uu = top_unit<UNITS>(pp, base0, ustep, inv_ustep);
                float r = 0.0f, ir = 0.0f;
                while (uu >= 0) {
                    const int k = pp - ((base0 + uu * ustep) & ~3);
                    if (k >= G.lb) break;
                    if constexpr (NF == 512) {
                        const float2 a = reinterpret_cast<const float2*>(pb)[uu * G.pitch + k];
                        r += a.x;
                        ir += a.y;
                    } else {
                        r += pb[uu * G.pitch + k];
                        ir += pb[(UNITS + uu) * G.pitch + k];
                    }
                    --uu;
                }
                r *= e;
                ir *= e;
                rrow[q] = r;
                irow[q] = ir;
                acc[0] += (double)r;
                acc[1] += (double)(r * r);
                acc[2] += (double)ir;
                acc[3] += (double)(ir * ir);
            }
        }
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

// ------------------------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------------------------
static int sm_count_dev() { return sm_count(); }
template <class K>
static int resident3(K kernel, int threads, size_t smem, int cap) { return resident_memo(kernel, threads, smem, cap); }

#ifdef ADV_AB
static bool stft3_vec() {
    static const char* e = ADV_AB_ENV("ADV_STFT3_VEC");  // A/B: 64-bit (default) or planar ("0") exchanges in the STFT
    static const bool on = !(e && e[0] == '0');
    return on;
}
#endif

template <int NF, bool RECT, bool VEC>
static int launch_stft3_t(const adv_plan* p, const float* wav, int64_t wav_stride, int batch, float2* X, float* mag,
                          float* phase, int flags, cudaStream_t s) {
    using C = S3Cfg<NF, VEC>;
    const bool zero_pad = (flags & ADV_STFT_ZERO_PAD) != 0;
    const size_t smem = C::bytes(p->d.hop, RECT);
    const int items_per_clip = (p->d.T + 1) / 2;
    const long total = (long)items_per_clip * batch;
    if (total > 0x7fffffffL) return ADV_ERR_UNSUPPORTED;
    const long ctas = (total + C::WARPS - 1) / C::WARPS;
    int rc;
#define ADV_LAUNCH_STFT3(M, PH, ZP)                                                                              \
    do {                                                                                                         \
        auto kernel = stft3_kernel<NF, M, PH, RECT, VEC, ZP>;                                                    \
        if ((rc = set_smem(kernel, smem)) != ADV_OK) return rc;                                                  \
        const long slots = (long)resident3(kernel, kThreads, smem, VEC ? ADV_STFT3_VEC_CTAS : 4) * sm_count_dev();                \
        const int grid = (int)(ctas < slots ? ctas : slots);                                                     \
        int groups = kWorkGroups;                                                                                \
        while (grid % groups != 0) --groups;                                                                     \
        ADV_CUDA_CHECK(launch_pdl(kernel, grid, kThreads, smem, s, p->d, wav, wav_stride, (int)total,            \
                                  items_per_clip, X, mag, phase, next_work_slot(p), groups));                    \
    } while (0)
    if (zero_pad) {  // the adjoint-of-istft use (training backward): spectrum only
        if (mag || phase) return ADV_ERR_UNSUPPORTED;
        ADV_LAUNCH_STFT3(false, false, true);
    } else if (mag && phase) ADV_LAUNCH_STFT3(true, true, false);
    else if (mag) ADV_LAUNCH_STFT3(true, false, false);
    else if (phase) ADV_LAUNCH_STFT3(false, true, false);
    else ADV_LAUNCH_STFT3(false, false, false);
#undef ADV_LAUNCH_STFT3
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

int launch_stft3(const adv_plan* p, const float* wav, int64_t wav_stride, int batch, float2* X, float* mag,
                 float* phase, int flags, cudaStream_t s) {
    if (!p->gen3 || !gen3_enabled()) return ADV_ERR_UNSUPPORTED;
#ifdef ADV_AB
    // n_fft 1024: the generation-2 warp-autonomous kernel (a warp per frame, 32 values per lane) measures faster than the
    // half-size complex transform with its twiddle pass here - 64 x 5 s clips, hop 322: X only 22.4 vs 27.4 us,
    // X + |X| + angle 39.5 vs 43.4 us (profiles/r02c_kbench_gen3_vs_gen2.jsonl).  ADV_STFT3_1024=1 keeps this kernel.
    static const char* e1024 = ADV_AB_ENV("ADV_STFT3_1024");
    if (p->d.n_fft == 1024 && !(e1024 && e1024[0] == '1')) return ADV_ERR_UNSUPPORTED;
    const bool vec = stft3_vec();   // ADV_STFT3_VEC=0: planar exchanges (measured slower: 18.9 vs 17.2 us X only)
    if (p->d.n_fft == 512) {
        if (p->d.rect_full)
            return vec ? launch_stft3_t<512, true, true>(p, wav, wav_stride, batch, X, mag, phase, flags, s)
                       : launch_stft3_t<512, true, false>(p, wav, wav_stride, batch, X, mag, phase, flags, s);
        return vec ? launch_stft3_t<512, false, true>(p, wav, wav_stride, batch, X, mag, phase, flags, s)
                   : launch_stft3_t<512, false, false>(p, wav, wav_stride, batch, X, mag, phase, flags, s);
    }
    // n_fft 1024: the window is always applied (a rectangular full-frame window multiplies by ones)
    return vec ? launch_stft3_t<1024, false, true>(p, wav, wav_stride, batch, X, mag, phase, flags, s)
               : launch_stft3_t<1024, false, false>(p, wav, wav_stride, batch, X, mag, phase, flags, s);
#else
    // product build: n_fft 512 with 64-bit exchanges; n_fft 1024 stays on the generation-2 warp-autonomous kernel, which
    // measures faster than the half-size complex transform here (22.4 vs 27.4 us X only, 64 x 5 s clips, hop 322)
    if (p->d.n_fft != 512) return ADV_ERR_UNSUPPORTED;
    return p->d.rect_full ? launch_stft3_t<512, true, true>(p, wav, wav_stride, batch, X, mag, phase, flags, s)
                          : launch_stft3_t<512, false, true>(p, wav, wav_stride, batch, X, mag, phase, flags, s);
#endif
}

int istft3_frames_cap(const adv_plan* p) {
    if (!p->gen3 || !gen3_enabled() || p->d.n_fft != 1024) return 0;
    if (p->max_hops_cap[1] <= 0) return 0;
    if (I3Cfg<1024>::bytes(p->d.hop, p->d.wlo, p->d.whi, false) > 113 * 1024) return 0;
    return 32;
}

template <int NF, bool RECT, int HS, bool CONTIG, int GV>
static int launch_istft3_t(const adv_plan* p, const Tiling& tl, const float2* X, int64_t sb, int64_t st, int64_t sf,
                           int batch, float* out, double* stats, cudaStream_t s) {
    const size_t smem = I3Cfg<NF>::bytes(p->d.hop, p->d.wlo, p->d.whi, RECT);
    const long total = (long)tl.tiles * batch;
    if (total > 0x7fffffffL) return ADV_ERR_UNSUPPORTED;
    auto kernel = istft3_kernel<NF, RECT, HS, CONTIG, GV>;
    int rc = set_smem(kernel, smem);
    if (rc != ADV_OK) return rc;
    const long slots = (long)resident3(kernel, kT3, smem, 2) * sm_count_dev();
    const int grid = (int)(total < slots ? total : slots);
    ADV_CUDA_CHECK(launch_pdl(kernel, grid, kT3, smem, s, p->d, tl, (int)total, X, sb, st, sf, out, stats));
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

int launch_istft3(const adv_plan* p, const float2* X, int64_t sb, int64_t st, int64_t sf, int batch, float* out,
                  double* stats, cudaStream_t s) {
    if (!p->gen3 || !gen3_enabled()) return ADV_ERR_UNSUPPORTED;
    const bool vec4 = p->d.n_out % 4 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0;
    if (p->d.n_fft == 512) {
        // same domain as the generation-2 wide kernel: contiguous rows, 4-sample groups aligned
        if (!(sf == 1 && p->d.hop % 4 == 0 && vec4) || I3Cfg<512>::bytes(p->d.hop, p->d.wlo, p->d.whi, false) > 110 * 1024)
            return ADV_ERR_UNSUPPORTED;
        const Tiling tl = choose_tiling(p, batch, 2, istft_balanced());
        if (p->d.rect_full && p->d.hop == 160) return launch_istft3_t<512, true, 5, true, 4>(p, tl, X, sb, st, sf, batch, out, stats, s);
        if (p->d.rect_full && p->d.hop == 128) return launch_istft3_t<512, true, 4, true, 4>(p, tl, X, sb, st, sf, batch, out, stats, s);
        if (p->d.rect_full && p->d.hop == 256) return launch_istft3_t<512, true, 8, true, 4>(p, tl, X, sb, st, sf, batch, out, stats, s);
        return p->d.rect_full ? launch_istft3_t<512, true, 0, true, 4>(p, tl, X, sb, st, sf, batch, out, stats, s)
                              : launch_istft3_t<512, false, 0, true, 4>(p, tl, X, sb, st, sf, batch, out, stats, s);
    }
    if (istft3_frames_cap(p) != 32) return ADV_ERR_UNSUPPORTED;
    Tiling tl = choose_tiling(p, batch, 2, istft_balanced(), 32);
    // the 128-bit gather needs every tile to start on a multiple of 4 samples
    const bool g4 = vec4 && ((long)tl.hops_per_tile * p->d.hop) % 4 == 0;
    if (sf == 1)
        return g4 ? launch_istft3_t<1024, false, 0, true, 4>(p, tl, X, sb, st, sf, batch, out, stats, s)
                  : launch_istft3_t<1024, false, 0, true, 1>(p, tl, X, sb, st, sf, batch, out, stats, s);
    return g4 ? launch_istft3_t<1024, false, 0, false, 4>(p, tl, X, sb, st, sf, batch, out, stats, s)
              : launch_istft3_t<1024, false, 0, false, 1>(p, tl, X, sb, st, sf, batch, out, stats, s);
}

template <int NF, int MODE, bool RECT, bool FROM_SPEC, int GV>
static int launch_explain3_t(const adv_plan* p, const Tiling& tl, const float* wav, int64_t wav_stride, const float2* X,
                             int64_t sb, int64_t st, int64_t sf, const float* mask, int Fm, int Tm, int drop, int batch,
                             float* rel, float* irr, double* stats, cudaStream_t s) {
    const size_t smem = E3Cfg<NF>::bytes(p->d.hop, p->d.wlo, p->d.whi, RECT, FROM_SPEC);
    const long total = (long)tl.tiles * batch;
    if (smem > 227 * 1024 || total > 0x7fffffffL) return ADV_ERR_UNSUPPORTED;
    auto kernel = explain3_kernel<NF, MODE, RECT, FROM_SPEC, GV>;
    int rc = set_smem(kernel, smem);
    if (rc != ADV_OK) return rc;
    const int n_sm = sm_count_dev();
    const int grid = (int)(total < n_sm ? total : n_sm);
    ADV_CUDA_CHECK(launch_pdl(kernel, grid, kT3, smem, s, p->d, tl, (int)total, wav, wav_stride, X, sb, st, sf, mask, Fm, Tm,
                              drop, rel, irr, stats));
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

int launch_explain3(const adv_plan* p, const float* wav, int64_t wav_stride, const float2* X, int64_t sb, int64_t st,
                    int64_t sf, const float* mask, int Fm, int Tm, int mode_flags, int batch, float* rel, float* irr,
                    double* stats, cudaStream_t s) {
    if (!p->gen3 || !gen3_enabled()) return ADV_ERR_UNSUPPORTED;
    const int mode = mode_flags & 0xff, drop = (mode_flags & ADV_MASK_DROP_OUTSIDE) ? 1 : 0;
    const bool spec = X != nullptr;
    const Tiling tl = choose_tiling(p, batch, 1);
    const bool vec4 = p->d.n_out % 4 == 0 && (reinterpret_cast<uintptr_t>(rel) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(irr) & 15) == 0 && ((long)tl.hops_per_tile * p->d.hop) % 4 == 0;
#define ADV_E3(NF, MODE, RECT, SPEC, GV) \
    return launch_explain3_t<NF, MODE, RECT, SPEC, GV>(p, tl, wav, wav_stride, X, sb, st, sf, mask, Fm, Tm, drop, batch, rel, irr, stats, s)
    if (p->d.n_fft == 512) {
        if (spec || !vec4 || p->d.hop % 4 != 0) return ADV_ERR_UNSUPPORTED;
        if (mode == ADV_MASK_LOG1P) { if (p->d.rect_full) ADV_E3(512, ADV_MASK_LOG1P, true, false, 4); else ADV_E3(512, ADV_MASK_LOG1P, false, false, 4); }
        else { if (p->d.rect_full) ADV_E3(512, ADV_MASK_LINEAR, true, false, 4); else ADV_E3(512, ADV_MASK_LINEAR, false, false, 4); }
    }
    if (mode == ADV_MASK_LOG1P) {
        if (spec) { if (vec4) ADV_E3(1024, ADV_MASK_LOG1P, false, true, 4); else ADV_E3(1024, ADV_MASK_LOG1P, false, true, 1); }
        else { if (vec4) ADV_E3(1024, ADV_MASK_LOG1P, false, false, 4); else ADV_E3(1024, ADV_MASK_LOG1P, false, false, 1); }
    } else {
        if (spec) { if (vec4) ADV_E3(1024, ADV_MASK_LINEAR, false, true, 4); else ADV_E3(1024, ADV_MASK_LINEAR, false, true, 1); }
        else { if (vec4) ADV_E3(1024, ADV_MASK_LINEAR, false, false, 4); else ADV_E3(1024, ADV_MASK_LINEAR, false, false, 1); }
    }
#undef ADV_E3
}

}  // namespace adv
