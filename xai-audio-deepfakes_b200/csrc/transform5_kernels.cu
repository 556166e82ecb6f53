// Streaming overlap-add kernels for n_fft 1024 (the reference's default geometry: AudioProcessor() = 1024 / hop 322 / a
// 644-tap rectangular window, audioprocessor.py:23-31; also hifigan.py's hann 1024 / hop 256): the warp pipeline of
// transform4_kernels.cu without its two restrictions (rectangular FULL window, hop a multiple of 32).
//
// Reference call sites: torch.istft (audioprocessor.py:123-129); STFT -> mask / (1 - mask) -> 2 x iSTFT
// (LMAC_metrics.py:136-157 log1p, loss_function.py:36-47 linear).
//
// What the generation-3 kernels pay at this size (tiles of 32 frames, private strips, two CTA-wide barriers per tile,
// a gather pass with an index search per 4 samples): iSTFT 37 - 41 us per 64 x 5 s clips (32 - 35 % of HBM peak), fused
// explain 123 - 125 us (11.5 %).  Here:
//   * unit = ONE frame = one 512-point complex transform of a warp (fft3.cuh real-1024 pre / post passes); the batch is one
//     list of units, CTA c owns a contiguous run and walks it 8 (iSTFT) / 16 (explain) units at a time, one per warp;
//   * a frame's windowed samples [wlo, whi) go to the warp's strip in shared memory (64-bit stores: lane l holds samples
//     2 l + 64 j, 2 l + 64 j + 1).  Its first `hop` samples are the unit's HEAD: final once the strips of the NB =
//     ceil(sup / hop) - 1 previous units are added at offsets k * hop.  The head is assembled 64 bits per lane from the
//     (1 + NB) strips, scaled by the reciprocal envelope and stored - no tile, no halo recompute (only the first NB units of
//     a run that does not begin a clip are transformed for their strips alone), no gather pass, no index search;
//   * hand-offs are mbarriers between the warps concerned (strip full / empty per unit slot): no CTA-wide barrier in
//     steady state;
//   * the next unit's spectrum row travels in registers while the current one is transformed.
// Domain: n_fft 1024; hop and n_out even (64-bit accesses stay aligned; the window support is widened to even bounds);
// 64 <= hop <= 512; support <= 4 hops.  Everything else stays on the generation-3 kernels.
// Normaliser statistics: as in transform4_kernels.cu (per-warp slots, fixed fold order).
#include "transform_common.cuh"
#include "fft3.cuh"
#include "stream_common.cuh"

namespace adv {

constexpr int kI5Units = 8, kI5Threads = 32 * kI5Units, kS5MaxRows = 8;   // head rows of 64 samples: hop <= 512

struct Geo5 {
    int hop, wlo, sup, nb;   // window support [wlo, wlo + sup), predecessors reaching into a head
    int strip;               // floats per strip (sup rounded up to a multiple of 4)
};

template <int UNITS>
struct S5Cfg {
    static size_t bytes(int sup, bool rect) {
        return al16(sizeof(float2) * f3::TW_TOTAL) + al16(sizeof(float) * UNITS * f3::Scr<true>::FLOATS) +
               al16(sizeof(float) * UNITS * ((sup + 3) & ~3)) + (rect ? 0 : al16(sizeof(float) * 1024)) +
               al16(sizeof(uint64_t) * 2 * UNITS);
    }
};

// spectrum row of frame `t` -> the lanes that own its bins (slot layout of fft3.cuh::r1024_pre)
template <bool CONTIG>
__device__ __forceinline__ void load_row1024(const float2* __restrict__ xp, int64_t sf, int l, int q1, bool live, float2* xk,
                                             float2* xm) {
    const float2 zero2 = make_float2(0.f, 0.f);
    const int64_t sfe = CONTIG ? 1 : sf;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        xk[i] = live ? __ldg(xp + (l + 64 * i) * sfe) : zero2;
        xm[i] = live ? __ldg(xp + (512 - l - 64 * i) * sfe) : zero2;
        xk[4 + i] = live ? __ldg(xp + (q1 + 64 * i) * sfe) : zero2;
        xm[4 + i] = live ? __ldg(xp + (512 - q1 - 64 * i) * sfe) : zero2;
    }
    xk[8] = (live && l == 0) ? __ldg(xp + 256 * sfe) : zero2;
    xm[8] = xk[8];
}

// windowed time samples of a frame (v[j] = samples 2 l + 64 j, + 1) -> the warp's strip, positions n - wlo
template <bool RECT>
__device__ __forceinline__ void strip_store(const float2* v, int l, const Geo5& G, const float* win_s, float* strip) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int k = 64 * j + 2 * l - G.wlo;
        if ((unsigned)k < (unsigned)G.sup) {
            float2 o = v[j];
            if (!RECT) {
                const float2 ww = *reinterpret_cast<const float2*>(win_s + 64 * j + 2 * l);
                o.x *= ww.x;
                o.y *= ww.y;
            }
            *reinterpret_cast<float2*>(strip + k) = o;
        }
    }
}

template <bool RECT, bool CONTIG>
__global__ void __launch_bounds__(kI5Threads, 2)
istft5_kernel(PlanDev P, Geo5 G, int upc, int total_units, const float2* __restrict__ X, int64_t sb, int64_t st, int64_t sf,
              float* __restrict__ out, double* __restrict__ stats, int slots) {
    constexpr int UNITS = kI5Units, NT = kI5Threads;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Carver cv{smem_raw};
    float2* tw_s = cv.take<float2>(f3::TW_TOTAL);
    float* scratch = cv.take<float>(UNITS * f3::Scr<true>::FLOATS);
    float* strips = cv.take<float>(UNITS * G.strip);
    float* win_s = RECT ? nullptr : cv.take<float>(1024);
    uint64_t* full = cv.take<uint64_t>(2 * UNITS);
    uint64_t* empty = full + UNITS;

    const int tid = threadIdx.x, l = tid & 31;
    const int w = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const long NG = gridDim.x;
    const int g_begin = (int)(((long)blockIdx.x * total_units) / NG);
    const int g_end = (int)(((long)(blockIdx.x + 1) * total_units) / NG);
    const int b0 = g_begin / upc, u0 = g_begin - b0 * upc;
    const int halo = u0 < G.nb ? u0 : G.nb;
    const int start = g_begin - halo;
    const int n_pass = (g_end - start + UNITS - 1) / UNITS;

    if (tid == 0)
        for (int i = 0; i < UNITS; ++i) {
            mbar_init(full + i, 1);
            mbar_init(empty + i, G.nb);
        }
    for (int i = tid; i < f3::TW_TOTAL / 2; i += NT) cp_async16(tw_s + 2 * i, P.tw3 + 2 * i);
    if (!RECT)
        for (int i = tid; i < 1024 / 4; i += NT) cp_async16(win_s + 4 * i, P.window + 4 * i);
    cp_async_commit();
    pdl_launch_dependents();
    cp_async_wait_all();
    __syncthreads();
    pdl_wait();

    if (stats != nullptr && g_end > g_begin) {   // zero this CTA's statistics slots (and, for a clip's first CTA, the unused ones)
        const int b_last = (g_end - 1) / upc;
        for (int b = b0; b <= b_last; ++b) {
            const long gb = (long)b * upc;
            const int c_first = (int)(((gb + 1) * NG - 1) / total_units);
            const int c_last = (int)(((gb + upc) * NG - 1) / total_units);
            double* row = stats + (size_t)b * slots * 2;
            if (l < 2) row[(((int)blockIdx.x - c_first) * UNITS + w) * 2 + l] = 0.0;
            if ((int)blockIdx.x == c_first)
                for (int i = (c_last - c_first + 1) * UNITS * 2 + tid; i < slots * 2; i += NT) row[i] = 0.0;
        }
    }

    float* my = scratch + w * f3::Scr<true>::FLOATS;
    float* strip_w = strips + w * G.strip;
    const int q1 = l == 0 ? 32 : 64 - l;

    UnitPos pos;
    {
        const int g = start + w;
        pos.b = g / upc;
        pos.u = g - pos.b * upc;
    }
    float acc[2] = {0.f, 0.f};
    int acc_b = -1;
    auto flush = [&]() {
        const double q0 = warp_sum((double)acc[0]), q1s = warp_sum((double)acc[1]);
        if (stats != nullptr && acc_b >= 0 && l == 0) {
            const int c_first = (int)((((long)acc_b * upc + 1) * NG - 1) / total_units);
            double* row = stats + ((size_t)acc_b * slots + ((int)blockIdx.x - c_first) * UNITS + w) * 2;
            row[0] = q0;
            row[1] = q1s;
        }
        acc[0] = acc[1] = 0.f;
    };
    float2 xk[9], xm[9];
    auto load_rows = [&](const UnitPos& q, bool live) {
        load_row1024<CONTIG>(X + (size_t)q.b * sb + (size_t)q.u * st, sf, l, q1, live && q.u < P.T, xk, xm);
    };
    load_rows(pos, start + w < g_end);

    for (int p = 0; p < n_pass; ++p) {
        const int g = start + p * UNITS + w;
        const bool active = g < g_end;
        if (!active) break;   // (last pass only: nobody waits for an arrival of a warp without a unit; see explain4_kernel)
        const bool is_out = active && g >= g_begin;
        const UnitPos cur = pos;
        pos.advance(UNITS, upc);

        // -- 1. the unit's spectrum row was requested a pass ago; build the transform input, request the next row
        float2 v[16];
        f3::r1024_pre(v, l, tw_s, xk, xm);
        load_rows(pos, p + 1 < n_pass && g + UNITS < g_end);
        // -- 2. inverse transform; windowed samples to the warp's strip once its previous readers are done
        f3::fft_inverse<true>(v, l, tw_s, my);
        if (p >= 1) mbar_wait(empty + w, (p - 1) & 1);
        strip_store<RECT>(v, l, G, win_s, strip_w);
        __syncwarp();
        if (l == 0) mbar_arrive(full + w);

        // -- 3. head = own strip [0, hop) + strips of the previous NB units at offsets k * hop
        const int s_base = cur.u * G.hop + G.wlo - 512 + 2 * l;   // output sample of this lane's first pair
        float2 head[kS5MaxRows], env[kS5MaxRows];
#pragma unroll
        for (int r = 0; r < kS5MaxRows; ++r) {
            const int i = 2 * l + 64 * r;
            const int s = s_base + 64 * r;
            const bool in = is_out && i < G.hop && s >= 0 && s < P.n_out;
            env[r] = in ? __ldg(reinterpret_cast<const float2*>(P.inv_env + s)) : make_float2(0.f, 0.f);
            head[r] = (i < G.hop && i < G.sup) ? *reinterpret_cast<const float2*>(strip_w + i) : make_float2(0.f, 0.f);
        }
        for (int k = G.nb; k >= 1; --k) {
            const int slot = (w - k) & (UNITS - 1);
            const int pp = w >= k ? p : p - 1;
            if (pp >= 0) mbar_wait(full + slot, pp & 1);
            if (is_out && cur.u >= k) {
                const float* tn = strips + slot * G.strip + k * G.hop;
#pragma unroll
                for (int r = 0; r < kS5MaxRows; ++r) {
                    const int i = 2 * l + 64 * r;
                    if (i < G.hop && i + k * G.hop < G.sup) {
                        const float2 t = *reinterpret_cast<const float2*>(tn + i);
                        head[r].x += t.x;
                        head[r].y += t.y;
                    }
                }
            }
            __syncwarp();
            if (l == 0 && pp >= 0) mbar_arrive(empty + slot);
        }
        // -- 4. scale, statistics, store
        if (is_out) {
            if (cur.b != acc_b) {
                if (acc_b >= 0) flush();
                acc_b = cur.b;
            }
            float* orow = out + (size_t)cur.b * P.n_out;
#pragma unroll
            for (int r = 0; r < kS5MaxRows; ++r) {
                const int i = 2 * l + 64 * r;
                const int s = s_base + 64 * r;
                if (i < G.hop && s >= 0 && s < P.n_out) {
                    const float2 a = make_float2(head[r].x * env[r].x, head[r].y * env[r].y);
                    acc[0] += a.x + a.y;
                    acc[1] = fmaf(a.x, a.x, fmaf(a.y, a.y, acc[1]));
                    *reinterpret_cast<float2*>(orow + s) = a;
                }
            }
        }
    }
    if (acc_b >= 0) flush();
}

// ------------------------------------------------------------------------------------------------------------------
// Fused explain, n_fft 1024: STFT -> mask / (1 - mask) -> 2 x iSTFT, one frame per warp and pass (16 warps).  Per unit: the
// frame's window support arrives as a per-warp bulk-async slice (reflect-padded clip edges by plain loads), one forward
// transform + real-1024 post pass gives bins k and 512 - k per slot, the mask column of the frame comes from a [513][17]
// tile staged by cp.async for the whole pass (one tile: the next one is requested once all 16 warps have read the
// current one), gains as in explain4_kernel, then two inverse transforms (masked-in, masked-out) whose windowed samples go
// to the warp's two strips; heads are assembled from the strips of the unit and its NB predecessors as in istft5_kernel.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kE5Units = 16, kE5Threads = 32 * kE5Units, kE5MP = kE5Units + 1, kE5F = 513;

struct E5Cfg {
    static constexpr int MASK_TILE = (kE5F * kE5MP + 3) & ~3;
    static __host__ __device__ int seg_floats(int sup) { return (sup + 8 + 3) & ~3; }
    static size_t bytes(int sup, bool rect) {
        return al16(sizeof(float2) * f3::TW_TOTAL) + al16(sizeof(float) * kE5Units * f3::Scr<false>::FLOATS) +
               al16(sizeof(float) * kE5Units * 2 * ((sup + 3) & ~3)) + al16(sizeof(float) * kE5Units * seg_floats(sup)) +
               al16(sizeof(float) * MASK_TILE) + (rect ? 0 : al16(sizeof(float) * 1024)) +
               al16(sizeof(uint64_t) * (3 * kE5Units + 2)) + al16(sizeof(float) * 4 * kE5Threads);
    }
};

template <int MODE, bool RECT>
__global__ void __maxnreg__(128)
explain5_kernel(PlanDev P, Geo5 G, int upc, int total_units, const float* __restrict__ wav, int64_t wav_stride,
                const float* __restrict__ mask, int Fm, int Tm, int drop, float* __restrict__ rel, float* __restrict__ irr,
                double* __restrict__ stats, int slots) {
    constexpr int UNITS = kE5Units, NT = kE5Threads, F = kE5F, MP = kE5MP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Carver cv{smem_raw};
    float2* tw_s = cv.take<float2>(f3::TW_TOTAL);
    float* scratch = cv.take<float>(UNITS * f3::Scr<false>::FLOATS);
    float* strips = cv.take<float>(UNITS * 2 * G.strip);
    const int SEG = E5Cfg::seg_floats(G.sup);
    float* seg_all = cv.take<float>(UNITS * SEG);
    float* mask_s = cv.take<float>(E5Cfg::MASK_TILE);
    float* win_s = RECT ? nullptr : cv.take<float>(1024);
    uint64_t* bars = cv.take<uint64_t>(3 * UNITS + 2);
    float* acc_s = cv.take<float>(4 * NT) + threadIdx.x;   // per-thread statistics accumulators [4][NT] (see explain4_kernel)
    uint64_t* full = bars;                 // [16] strips of unit slot w written
    uint64_t* empty = bars + UNITS;        // [16] strips of unit slot w read by their NB consumers
    uint64_t* segbar = bars + 2 * UNITS;   // [16] waveform slice of warp w landed
    uint64_t* mfull = bars + 3 * UNITS;    // mask tile landed (512 cp.async arrivals per pass)
    uint64_t* mempty = mfull + 1;          // mask tile read by all 16 warps

    const int tid = threadIdx.x, l = tid & 31;
    const int w = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const long NG = gridDim.x;
    const int g_begin = (int)(((long)blockIdx.x * total_units) / NG);
    const int g_end = (int)(((long)(blockIdx.x + 1) * total_units) / NG);
    const int b0 = g_begin / upc, u0 = g_begin - b0 * upc;
    const int halo = u0 < G.nb ? u0 : G.nb;
    const int start = g_begin - halo;
    const int n_pass = (g_end - start + UNITS - 1) / UNITS;
    const int T_eff = drop ? (Tm < P.T ? Tm : P.T) : P.T;   // frames past the mask are dropped from both outputs
    const int f_lim = drop ? Fm : F;

    if (tid == 0) {
        for (int i = 0; i < UNITS; ++i) {
            mbar_init(full + i, 1);
            mbar_init(empty + i, G.nb);
            mbar_init(segbar + i, 1);
        }
        mbar_init(mfull, NT);
        mbar_init(mempty, UNITS);
    }
    for (int i = tid; i < f3::TW_TOTAL / 2; i += NT) cp_async16(tw_s + 2 * i, P.tw3 + 2 * i);
    if (!RECT)
        for (int i = tid; i < 1024 / 4; i += NT) cp_async16(win_s + 4 * i, P.window + 4 * i);
    cp_async_commit();
    pdl_launch_dependents();
    cp_async_wait_all();
    __syncthreads();  // tables staged, barriers initialised (the only CTA-wide barrier of the kernel)
    pdl_wait();

    if (stats != nullptr && g_end > g_begin) {
        const int b_last = (g_end - 1) / upc;
        for (int b = b0; b <= b_last; ++b) {
            const long gb = (long)b * upc;
            const int c_first = (int)(((gb + 1) * NG - 1) / total_units);
            const int c_last = (int)(((gb + upc) * NG - 1) / total_units);
            double* row = stats + (size_t)b * slots * 4;
            if (l < 4) row[(((int)blockIdx.x - c_first) * UNITS + w) * 4 + l] = 0.0;
            if ((int)blockIdx.x == c_first)
                for (int i = (c_last - c_first + 1) * UNITS * 4 + tid; i < slots * 4; i += NT) row[i] = 0.0;
        }
    }

    float* my = scratch + w * f3::Scr<false>::FLOATS;
    float* seg = seg_all + w * SEG;
    float* strip_r = strips + w * (2 * G.strip);
    float* strip_i = strip_r + G.strip;
    const int q1 = l == 0 ? 32 : 64 - l;

    // ---- per-thread mask staging: column c of the tile <-> unit slot c; rows f0 + 32 k
    const int mc = tid & 15, mf0 = tid >> 4;
    UnitPos mpos;
    {
        const int g = start + mc;
        mpos.b = g / upc;
        mpos.u = g - mpos.b * upc;
    }
    auto request_mask = [&](int pass) {
        const int g = start + pass * UNITS + mc;
        const int t = mpos.u;
        const bool col_ok = g < g_end && t < Tm;
        const float* mrow = mask + (size_t)mpos.b * Fm * Tm;
        const float* src = col_ok ? mrow + (size_t)mf0 * Tm + t : mask;
        const size_t step = col_ok ? (size_t)32 * Tm : 0;
        uint32_t dst = smem_u32(mask_s + mf0 * MP + mc);
        const int full_rows = col_ok ? (Fm - mf0 + 31) / 32 : 0;   // trips whose row exists in the mask
#pragma unroll
        for (int k = 0; k < 17; ++k) {
            if (mf0 + 32 * k < F) {
                const int sz = k < full_rows ? 4 : 0;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + k * 32 * MP * 4), "l"(src), "r"(sz) : "memory");
            }
            src += step;
        }
        cp_async_arrive_noinc(mfull);
        mpos.advance(UNITS, upc);
    };

    UnitPos pos;
    {
        const int g = start + w;
        pos.b = g / upc;
        pos.u = g - pos.b * upc;
    }
    auto request_seg = [&](const UnitPos& q, bool live) -> int {
        const int base = q.u * G.hop - 512 + G.wlo;
        if (live) return stage_segment_async<32, true>(seg, G.sup, wav + (size_t)q.b * wav_stride, base, P.n_in, segbar + w, l);
        if (l == 0) mbar_expect_tx(segbar + w, 0);   // (a unit past the run still arms the barrier)
        return 0;
    };

    request_mask(0);
    int shift = request_seg(pos, start + w < g_end);

#pragma unroll
    for (int i = 0; i < 4; ++i) acc_s[i * NT] = 0.f;
    int acc_b = -1;
    auto flush = [&]() {
        double q[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) q[i] = warp_sum((double)acc_s[i * NT]);
        if (stats != nullptr && acc_b >= 0 && l == 0) {
            const int c_first = (int)((((long)acc_b * upc + 1) * NG - 1) / total_units);
            double* row = stats + ((size_t)acc_b * slots + ((int)blockIdx.x - c_first) * UNITS + w) * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) row[i] = q[i];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) acc_s[i * NT] = 0.f;
    };

    for (int p = 0; p < n_pass; ++p) {
        const int g = start + p * UNITS + w;
        const bool active = g < g_end;
        if (!active) break;   // (last pass only: nobody waits for an arrival of a warp without a unit; see explain4_kernel)
        const bool is_out = active && g >= g_begin;
        const UnitPos cur = pos;
        const int cur_shift = shift;

        // -- 1. windowed samples of the frame (zeros outside the window support, for frames that do not exist or are
        //       dropped, and for units past the run - no branch encloses the transforms)
        float2 v[16];
        cp_async_wait_group<0>();   // the slice's edge samples (groups committed so far: not the mask copies requested after them)
        __syncwarp();
        mbar_wait(segbar + w, p & 1);
        {
            const bool live = active && cur.u < T_eff;
            const float* sp = seg + cur_shift + 2 * l - G.wlo;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int k = 64 * j + 2 * l - G.wlo;
                float2 x = make_float2(0.f, 0.f);
                if (live && (unsigned)k < (unsigned)G.sup) {
                    x = *reinterpret_cast<const float2*>(sp + 64 * j);
                    if (!RECT) {
                        const float2 ww = *reinterpret_cast<const float2*>(win_s + 64 * j + 2 * l);
                        x.x *= ww.x;
                        x.y *= ww.y;
                    }
                }
                v[j] = x;
            }
        }
        pos.advance(UNITS, upc);

        float2 zk[9], zm[9];   // masked-out spectrum, kept across the first inverse transform
        {
            // -- 2. forward transform (the next pass's slice is requested once every lane has consumed its samples)
            f3::fft_forward<false>(v, l, tw_s, my, [&] { if (p + 1 < n_pass) shift = request_seg(pos, g + UNITS < g_end); });
            float2 xk[9], xm[9];
            f3::r1024_post(v, l, tw_s, xk, xm);
            if (f_lim < F) {   // (uniform) outside="drop" with a mask narrower than the spectrum
#pragma unroll
                for (int i = 0; i < 9; ++i) {
                    const int k = i < 4 ? l + 64 * i : (i < 8 ? q1 + 64 * (i - 4) : 256);
                    if (k >= f_lim) xk[i] = make_float2(0.f, 0.f);
                    if (512 - k >= f_lim) xm[i] = make_float2(0.f, 0.f);
                }
            }
            // -- 3. mask column of the frame (bins k and 512 - k per slot), gains
            mbar_wait(mfull, p & 1);
            float mk[9], mm[9];
            {
                const float* c = mask_s + w;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    mk[i] = c[(l + 64 * i) * MP];
                    mm[i] = c[(512 - l - 64 * i) * MP];
                    mk[4 + i] = c[(q1 + 64 * i) * MP];
                    mm[4 + i] = c[(512 - q1 - 64 * i) * MP];
                }
                mk[8] = c[256 * MP];
                mm[8] = mk[8];
            }
            float2 yk[9], ym[9];
            frame_gains<MODE>(xk, mk, yk, zk);
            frame_gains<MODE>(xm, mm, ym, zm);
            __syncwarp();
            if (l == 0) mbar_arrive(mempty);   // (the gains consumed the mask values: the loads have returned)
            // -- 4. masked-in inverse transform -> the warp's first strip
            f3::r1024_pre(v, l, tw_s, yk, ym);
        }
        f3::fft_inverse<false>(v, l, tw_s, my);
        if (p >= 1) mbar_wait(empty + w, (p - 1) & 1);   // the previous pass's strips of this slot have been consumed
        strip_store<RECT>(v, l, G, win_s, strip_r);
        // -- 5. masked-out inverse transform -> the second strip
        f3::r1024_pre(v, l, tw_s, zk, zm);
        f3::fft_inverse<false>(v, l, tw_s, my);
        strip_store<RECT>(v, l, G, win_s, strip_i);
        __syncwarp();
        if (l == 0) mbar_arrive(full + w);
        // -- this thread's share of the next pass's mask tile, once every warp has read the current one
        if (p + 1 < n_pass) {
            mbar_wait(mempty, p & 1);
            request_mask(p + 1);
        }

        // -- 6. heads: own strips [0, hop) + strips of the previous NB units at offsets k * hop
        const int s_base = cur.u * G.hop + G.wlo - 512 + 2 * l;
        float2 hr[kS5MaxRows], hi[kS5MaxRows], env[kS5MaxRows];
#pragma unroll
        for (int r = 0; r < kS5MaxRows; ++r) {
            const int i = 2 * l + 64 * r;
            const int s = s_base + 64 * r;
            const bool in = is_out && i < G.hop && s >= 0 && s < P.n_out;
            env[r] = in ? __ldg(reinterpret_cast<const float2*>(P.inv_env + s)) : make_float2(0.f, 0.f);
            const bool own = i < G.hop && i < G.sup;
            hr[r] = own ? *reinterpret_cast<const float2*>(strip_r + i) : make_float2(0.f, 0.f);
            hi[r] = own ? *reinterpret_cast<const float2*>(strip_i + i) : make_float2(0.f, 0.f);
        }
        for (int k = G.nb; k >= 1; --k) {
            const int slot = (w - k) & (UNITS - 1);
            const int pp = w >= k ? p : p - 1;
            if (pp >= 0) mbar_wait(full + slot, pp & 1);
            if (is_out && cur.u >= k) {
                const float* tr = strips + slot * (2 * G.strip) + k * G.hop;
                const float* ti = tr + G.strip;
#pragma unroll
                for (int r = 0; r < kS5MaxRows; ++r) {
                    const int i = 2 * l + 64 * r;
                    if (i < G.hop && i + k * G.hop < G.sup) {
                        const float2 a = *reinterpret_cast<const float2*>(tr + i);
                        const float2 c = *reinterpret_cast<const float2*>(ti + i);
                        hr[r].x += a.x; hr[r].y += a.y;
                        hi[r].x += c.x; hi[r].y += c.y;
                    }
                }
            }
            __syncwarp();
            if (l == 0 && pp >= 0) mbar_arrive(empty + slot);
        }
        // -- 7. scale, statistics, store
        if (is_out) {
            if (cur.b != acc_b) {
                if (acc_b >= 0) flush();
                acc_b = cur.b;
            }
            float* rrow = rel + (size_t)cur.b * P.n_out;
            float* irow = irr + (size_t)cur.b * P.n_out;
            float acc[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = acc_s[i * NT];
#pragma unroll
            for (int r = 0; r < kS5MaxRows; ++r) {
                const int i = 2 * l + 64 * r;
                const int s = s_base + 64 * r;
                if (i < G.hop && s >= 0 && s < P.n_out) {
                    const float2 a = make_float2(hr[r].x * env[r].x, hr[r].y * env[r].y);
                    const float2 c = make_float2(hi[r].x * env[r].x, hi[r].y * env[r].y);
                    acc[0] += a.x + a.y;
                    acc[1] = fmaf(a.x, a.x, fmaf(a.y, a.y, acc[1]));
                    acc[2] += c.x + c.y;
                    acc[3] = fmaf(c.x, c.x, fmaf(c.y, c.y, acc[3]));
                    *reinterpret_cast<float2*>(rrow + s) = a;
                    *reinterpret_cast<float2*>(irow + s) = c;
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) acc_s[i * NT] = acc[i];
        }
    }
    if (acc_b >= 0) flush();
}

// ------------------------------------------------------------------------------------------------------------------
// STFT, n_fft 1024 (torch.stft + abs + angle, audioprocessor.py:102-110): warp-autonomous like stft3_kernel, but an item
// is ONE frame and its slice is the window support only (644 of 1024 samples for the reference's default window - the
// rest of the frame multiplies zeros), so the per-warp staging shrinks from hop + 1024 to ~650 floats and three 256-thread
// CTAs share an SM (the generation-3 form of this size held two, the generation-2 form 16 warps of 128 registers).
// A rectangular-on-its-support window costs no multiplies.  Reflect padding only (the zero-padded adjoint form stays
// on stft_w_kernel).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kF5Warps = 8, kF5Threads = 32 * kF5Warps;

struct F5Cfg {
    static __host__ __device__ int seg_floats(int sup) { return (sup + 8 + 3) & ~3; }
    static size_t bytes(int sup, bool rect) {
        return al16(sizeof(float2) * f3::TW_TOTAL) + al16(sizeof(float) * kF5Warps * f3::Scr<true>::FLOATS) +
               al16(sizeof(float) * kF5Warps * seg_floats(sup)) + (rect ? 0 : al16(sizeof(float) * 1024)) +
               al16(sizeof(uint64_t) * kF5Warps);
    }
};

template <bool MAG, bool PHASE>
__device__ __forceinline__ void store_bin5(float2* __restrict__ X, float* __restrict__ mag, float* __restrict__ phase,
                                           size_t idx, float2 x) {
    X[idx] = x;
    if (MAG) mag[idx] = fast_abs2(x);
    if (PHASE) phase[idx] = fast_atan2f(x.y, x.x);
}

template <bool MAG, bool PHASE, bool RECT>
__global__ void __launch_bounds__(kF5Threads, 3)
stft5_kernel(PlanDev P, Geo5 G, const float* __restrict__ wav, int64_t wav_stride, int total_items,
             float2* __restrict__ X, float* __restrict__ mag, float* __restrict__ phase) {
    constexpr int F = 513, WARPS = kF5Warps;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Carver cv{smem_raw};
    float2* tw_s = cv.take<float2>(f3::TW_TOTAL);
    float* scratch = cv.take<float>(WARPS * f3::Scr<true>::FLOATS);
    const int SEG = F5Cfg::seg_floats(G.sup);
    float* seg_all = cv.take<float>(WARPS * SEG);
    float* win_s = RECT ? nullptr : cv.take<float>(1024);
    uint64_t* bars = cv.take<uint64_t>(WARPS);

    const int tid = threadIdx.x, l = tid & 31;
    const int w = __shfl_sync(0xffffffffu, tid >> 5, 0);   // (tells the compiler the warp index is warp-uniform)
    float* seg = seg_all + (size_t)w * SEG;
    uint64_t* bar = bars + w;
    const int stride = gridDim.x * WARPS;
    // per-warp trip count: a warp leaves the loop after its last item (it requests no slice it will not consume)
    const int first = blockIdx.x * WARPS + w;
    const int n_iter = first < total_items ? (total_items - first + stride - 1) / stride : 0;

    if (l == 0) mbar_init(bar, 1);
    for (int i = tid; i < f3::TW_TOTAL / 2; i += kF5Threads) cp_async16(tw_s + 2 * i, P.tw3 + 2 * i);
    if (!RECT)
        for (int i = tid; i < 1024 / 4; i += kF5Threads) cp_async16(win_s + 4 * i, P.window + 4 * i);
    cp_async_commit();
    pdl_launch_dependents();
    cp_async_wait_all();
    __syncthreads();
    pdl_wait();

    float* my = scratch + w * f3::Scr<true>::FLOATS;
    const int q1 = l == 0 ? 32 : 64 - l;
    int item = min(first, total_items - 1);
    int b = item / P.T, t = item - b * P.T;
    int shift = 0;
    if (n_iter > 0)
        shift = stage_segment_async<32>(seg, G.sup, wav + (size_t)b * wav_stride, t * G.hop - 512 + G.wlo, P.n_in, bar, l);

    for (int it = 0; it < n_iter; ++it) {
        __syncwarp();
        mbar_wait(bar, it & 1);
        const int cur_b = b, cur_t = t, cur_shift = shift;
        constexpr bool active = true;
        const bool more = it + 1 < n_iter;
        float2 v[16];
        {
            const float* sp = seg + cur_shift + 2 * l - G.wlo;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int k = 64 * j + 2 * l - G.wlo;
                float2 x = make_float2(0.f, 0.f);
                if ((unsigned)k < (unsigned)G.sup) {
                    x = *reinterpret_cast<const float2*>(sp + 64 * j);
                    if (!RECT) {
                        const float2 ww = *reinterpret_cast<const float2*>(win_s + 64 * j + 2 * l);
                        x.x *= ww.x;
                        x.y *= ww.y;
                    }
                }
                v[j] = x;
            }
        }
        f3::fft_forward<true>(v, l, tw_s, my, [&] {   // every lane has consumed its samples: request the next slice
            if (more) {
                item = first + (it + 1) * stride;
                b = item / P.T;
                t = item - b * P.T;
                shift = stage_segment_async<32>(seg, G.sup, wav + (size_t)b * wav_stride, t * G.hop - 512 + G.wlo, P.n_in,
                                                bar, l);
            }
        });
        float2 xk[9], xm[9];
        f3::r1024_post(v, l, tw_s, xk, xm);
        if (active) {
            const size_t row = ((size_t)cur_b * P.T + cur_t) * F;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                store_bin5<MAG, PHASE>(X, mag, phase, row + l + 64 * i, xk[i]);
                store_bin5<MAG, PHASE>(X, mag, phase, row + 512 - l - 64 * i, xm[i]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                store_bin5<MAG, PHASE>(X, mag, phase, row + q1 + 64 * i, xk[4 + i]);
                store_bin5<MAG, PHASE>(X, mag, phase, row + 512 - q1 - 64 * i, xm[4 + i]);
            }
            if (l == 0) store_bin5<MAG, PHASE>(X, mag, phase, row + 256, xk[8]);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
struct Run5 { int ok, upc, grid, slots, rect; long total; Geo5 g; };
static Run5 plan_run5(const adv_plan* p, int batch, int units, int ctas_per_sm) {
    Run5 r = {};
    const PlanDev& d = p->d;
    if (d.n_fft != 1024 || d.n_out <= 0 || batch <= 0) return r;
    // support widened to even bounds (a periodic hann window starts with an exact zero: wlo = 1); the extra taps are zeros
    // of the window table, so such a plan always multiplies by the window
    const int wlo = d.wlo & ~1, whi = (d.whi + 1) & ~1, sup = whi - wlo;
    r.rect = p->rect_sup && wlo == d.wlo && whi == d.whi;
    if ((d.hop | d.n_out) & 1) return r;
    if (d.hop < 64 || d.hop > 64 * kS5MaxRows || sup > 4 * d.hop || wlo > 512) return r;
    r.g.hop = d.hop; r.g.wlo = wlo; r.g.sup = sup;
    r.g.nb = (sup + d.hop - 1) / d.hop - 1;
    if (r.g.nb < 1) r.g.nb = 1;   // (hop >= support: no overlap; the protocol still wants one reader per strip)
    r.g.strip = (sup + 3) & ~3;
    r.upc = (d.n_out - 1 + 512 - wlo) / d.hop + 1;   // units whose heads cover every output sample
    r.total = (long)r.upc * batch;
    if (r.total > 0x3fffffffL) return r;
    const long by_work = r.total / 8 > 0 ? r.total / 8 : 1;   // at least 8 units per CTA
    const long slots_hw = (long)ctas_per_sm * sm_count();
    r.grid = (int)(by_work < slots_hw ? by_work : slots_hw);
    const long span = ((long)(r.upc - 1) * r.grid + r.total - 1) / r.total;   // CTAs that can touch one clip
    r.slots = units * (int)(span + 1);
    r.ok = 1;
    return r;
}

int istft5_slots(const adv_plan* p, int batch) {
    const Run5 r = plan_run5(p, batch, kI5Units, 2);
    return r.ok ? r.slots : 0;
}

template <bool RECT, bool CONTIG>
static int launch_istft5_t(const adv_plan* p, const Run5& r, const float2* X, int64_t sb, int64_t st, int64_t sf, float* out,
                           double* stats, cudaStream_t s) {
    const size_t smem = S5Cfg<kI5Units>::bytes(r.g.sup, RECT);
    auto kernel = istft5_kernel<RECT, CONTIG>;
    int rc = set_smem(kernel, smem);
    if (rc != ADV_OK) return rc;
    ADV_CUDA_CHECK(launch_pdl(kernel, r.grid, kI5Threads, smem, s, p->d, r.g, r.upc, (int)r.total, X, sb, st, sf, out, stats,
                              r.slots));
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

int launch_istft5(const adv_plan* p, const float2* X, int64_t sb, int64_t st, int64_t sf, int batch, float* out,
                  double* stats, cudaStream_t s) {
    const Run5 r = plan_run5(p, batch, kI5Units, 2);
    if (!r.ok || (reinterpret_cast<uintptr_t>(out) & 7) != 0) return ADV_ERR_UNSUPPORTED;
    const bool rect = r.rect != 0;
    if (sf == 1)
        return rect ? launch_istft5_t<true, true>(p, r, X, sb, st, sf, out, stats, s)
                    : launch_istft5_t<false, true>(p, r, X, sb, st, sf, out, stats, s);
    return rect ? launch_istft5_t<true, false>(p, r, X, sb, st, sf, out, stats, s)
                : launch_istft5_t<false, false>(p, r, X, sb, st, sf, out, stats, s);
}


// ---- STFT
template <bool RECT>
static int launch_stft5_t(const adv_plan* p, const Geo5& g, const float* wav, int64_t wav_stride, int batch, float2* X,
                          float* mag, float* phase, cudaStream_t s) {
    const size_t smem = F5Cfg::bytes(g.sup, RECT);
    const long total = (long)p->d.T * batch;
    if (total > 0x7fffffffL) return ADV_ERR_UNSUPPORTED;
    const long ctas = (total + kF5Warps - 1) / kF5Warps;
    int rc;
#define ADV_LAUNCH_STFT5(M, PH)                                                                                  \
    do {                                                                                                         \
        auto kernel = stft5_kernel<M, PH, RECT>;                                                                 \
        if ((rc = set_smem(kernel, smem)) != ADV_OK) return rc;                                                  \
        const long slots = (long)resident_memo(kernel, kF5Threads, smem, 3) * sm_count();                        \
        const int grid = (int)(ctas < slots ? ctas : slots);                                                     \
        ADV_CUDA_CHECK(launch_pdl(kernel, grid, kF5Threads, smem, s, p->d, g, wav, wav_stride, (int)total, X, mag, phase)); \
    } while (0)
    if (mag && phase) ADV_LAUNCH_STFT5(true, true);
    else if (mag) ADV_LAUNCH_STFT5(true, false);
    else if (phase) ADV_LAUNCH_STFT5(false, true);
    else ADV_LAUNCH_STFT5(false, false);
#undef ADV_LAUNCH_STFT5
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

int launch_stft5(const adv_plan* p, const float* wav, int64_t wav_stride, int batch, float2* X, float* mag, float* phase,
                 int flags, cudaStream_t s) {
    const PlanDev& d = p->d;
    if (d.n_fft != 1024 || d.n_in <= 0 || (flags & ADV_STFT_ZERO_PAD) || (d.hop & 1)) return ADV_ERR_UNSUPPORTED;
    Geo5 g = {};
    const int wlo = d.wlo & ~1, whi = (d.whi + 1) & ~1;
    g.hop = d.hop; g.wlo = wlo; g.sup = whi - wlo; g.nb = 0; g.strip = 0;
    const bool rect = p->rect_sup && wlo == d.wlo && whi == d.whi;
    // Measured (profiles/r02z_kbench_stft5.jsonl, 64 x 5 s clips, hop 322, 644-tap rectangular window): X only 22.4 -> 20.8 us
    // (58 -> 63 % of HBM peak; 70 % at 256 clips); hann 1024 / hop 256 X only 25.7 -> 23.9 us, X + |X| + angle 44.1 -> 38.4 us.
    // With |X| and angle on the rectangular default window the two-frames-per-transform generation-2 kernel needs fewer
    // instructions per frame (this kernel: 1 484, 63 % of the issue slots, ncu) and stays ahead - 40.4 vs 41.3 us, and 52.9 vs
    // 62.4 us on 16 x 30 s clips - so that call keeps it.
#ifndef ADV_STFT5_RECT_MP   // 1: also serve the rectangular window with magnitude / phase outputs (A/B against the generation-2 kernel)
#define ADV_STFT5_RECT_MP 0
#endif
    if (rect && (mag || phase) && !ADV_STFT5_RECT_MP) return ADV_ERR_UNSUPPORTED;
    return rect ? launch_stft5_t<true>(p, g, wav, wav_stride, batch, X, mag, phase, s)
                : launch_stft5_t<false>(p, g, wav, wav_stride, batch, X, mag, phase, s);
}

// ---- fused explain
static Run5 plan_run_e5(const adv_plan* p, int batch) {
    Run5 r = plan_run5(p, batch, kE5Units, 1);
    if (!r.ok) return r;
    if (p->d.n_in <= 0 || E5Cfg::bytes(r.g.sup, r.rect != 0) > 227 * 1024) r.ok = 0;   // wide windows: strips + slices do not fit
    return r;
}

int explain5_slots(const adv_plan* p, int batch) {
    const Run5 r = plan_run_e5(p, batch);
    return r.ok ? r.slots : 0;
}

template <int MODE, bool RECT>
static int launch_explain5_t(const adv_plan* p, const Run5& r, const float* wav, int64_t wav_stride, const float* mask, int Fm,
                             int Tm, int drop, float* rel, float* irr, double* stats, cudaStream_t s) {
    const size_t smem = E5Cfg::bytes(r.g.sup, RECT);
    auto kernel = explain5_kernel<MODE, RECT>;
    int rc = set_smem(kernel, smem);
    if (rc != ADV_OK) return rc;
    ADV_CUDA_CHECK(launch_pdl(kernel, r.grid, kE5Threads, smem, s, p->d, r.g, r.upc, (int)r.total, wav, wav_stride, mask, Fm, Tm,
                              drop, rel, irr, stats, r.slots));
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

int launch_explain5(const adv_plan* p, const float* wav, int64_t wav_stride, const float* mask, int Fm, int Tm,
                    int mode_flags, int batch, float* rel, float* irr, double* stats, cudaStream_t s) {
    const Run5 r = plan_run_e5(p, batch);
    if (!r.ok || wav == nullptr) return ADV_ERR_UNSUPPORTED;
    if (((reinterpret_cast<uintptr_t>(rel) | reinterpret_cast<uintptr_t>(irr)) & 7) != 0) return ADV_ERR_UNSUPPORTED;
    const int mode = mode_flags & 0xff, drop = (mode_flags & ADV_MASK_DROP_OUTSIDE) ? 1 : 0;
    if (mode == ADV_MASK_LOG1P)
        return r.rect ? launch_explain5_t<ADV_MASK_LOG1P, true>(p, r, wav, wav_stride, mask, Fm, Tm, drop, rel, irr, stats, s)
                      : launch_explain5_t<ADV_MASK_LOG1P, false>(p, r, wav, wav_stride, mask, Fm, Tm, drop, rel, irr, stats, s);
    return r.rect ? launch_explain5_t<ADV_MASK_LINEAR, true>(p, r, wav, wav_stride, mask, Fm, Tm, drop, rel, irr, stats, s)
                  : launch_explain5_t<ADV_MASK_LINEAR, false>(p, r, wav, wav_stride, mask, Fm, Tm, drop, rel, irr, stats, s);
}

}  // namespace adv
