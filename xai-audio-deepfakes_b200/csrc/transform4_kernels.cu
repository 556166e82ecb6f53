// Generation-4 fused explain kernel for sm_100a: STFT -> mask / (1 - mask) -> 2 x iSTFT as a STREAMING, barrier-free
// pipeline of warps (reference call sites: LMAC_metrics.py:136-157 log1p, loss_function.py:36-47 linear).
//
// What the ncu captures of the generation-3 kernel (profiles/r02c_*) say bounds it: 50.3 M warp-instructions per
// 64 clips at 59 % issue utilisation while an SM is busy, and SMs busy only 80 % of the launch:
//   * tiles of 32 frames deliver 29 hops (halo frames are transformed twice) and 896 tiles on 148 SMs bill 7 rounds
//     for 6.05 rounds of work;
//   * two CTA-wide barriers per tile keep the 16 warps of an SM in lock-step phases (LDS/STS bursts, then FMA bursts);
//   * 39 % of the instructions are FFT work; gains (28 %), the strip gather (13 %) and mask / segment staging and
//     index arithmetic (20 %) are the rest.
// Here (n_fft 512, rectangular full window, hop = 32 HS: the benchmark geometry):
//   * the batch is ONE list of units (unit = two adjacent frames of a clip); CTA c owns a contiguous run of it and walks
//     it in passes of 16 units, one unit per warp.  No tile halo: a unit's overlap into later samples (its "tail",
//     16 - HS rows of 32 samples) is handed to the next NB units through shared memory; only the first NB units of a
//     run that does not begin a clip are transformed for their tails alone (148 x 2 units per launch instead of 10 %);
//   * overlap-add is register-local: inverse transform 1 carries the masked-in waveforms of frames a and b as real /
//     imaginary part, transform 2 the masked-out ones; frame b's sample 32 j + l lands on strip row j + HS of the SAME
//     lane.  A unit's first 2 HS rows ("head") are final once the tails of units u-1 .. u-NB are added - they go from
//     registers straight to global memory (128-byte rows); no strips, no gather pass, no index search;
//   * every hand-off is an mbarrier between the warps concerned (tail full / empty per unit slot, per-warp waveform
//     slices by bulk-async copy, the double-buffered mask tile by cp.async + mbarrier.arrive.noinc): there is NO
//     CTA-wide barrier in steady state and the warps drift apart, so one warp's shared-memory exchange overlaps
//     another's butterflies;
//   * gains: MUFU .approx.ftz forms without the denormal fix-ups, the small-magnitude series only where a frame has
//     such a bin (warp vote), out-of-mask handling hoisted out of the bin loop.
// Normaliser statistics: every warp keeps (sum, sum of squares) of what it stores per clip and writes them to slot
// (CTA's rank among the CTAs touching the clip) * 16 + warp of stats[B][slots][4]; unused slots are zeroed, the fold
// order of the normaliser stays fixed (bit-reproducible run to run).
#include "transform_common.cuh"
#include "fft3.cuh"
#include "stream_common.cuh"

namespace adv {

#ifndef ADV_E4_MASK_GW
#define ADV_E4_MASK_GW 16
#endif
#ifndef ADV_E4_LATE_EMPTY
#define ADV_E4_LATE_EMPTY 1
#endif
#ifndef ADV_E4_EARLY_EXIT  // 1: warps without a unit in the (last) pass leave the loop instead of transforming silence
#define ADV_E4_EARLY_EXIT 1
#endif
template <int HS>
struct E4Cfg {
    static constexpr int UNITS = 16, NT = 512, F = 257, MP = 33;
    static constexpr int HOP = 32 * HS, USTEP = 2 * HOP;
    static constexpr int ROWS = 16 + HS;                  // strip rows (32 samples each) of a unit: frame a, frame b HS rows later
    static constexpr int HEAD = 2 * HS;                   // rows a unit owns (its two hops)
    static constexpr int TAIL = ROWS - HEAD;              // rows handed to later units
    static constexpr int NB = (TAIL + HEAD - 1) / HEAD;   // earlier units reaching into a unit's head
    static constexpr int SEG = (HOP + 512 + 8 + 3) & ~3;  // floats of a warp's waveform slice
    static constexpr int MASK_TILE = (F * MP + 3) & ~3;
    // warps sharing one refill of their mask columns (16 = the whole tile at once; see ADV_E4_MASK_GW below)
    static constexpr int MGW = ADV_E4_MASK_GW, MGROUPS = UNITS / MGW;
    static constexpr int NBARS = 3 * UNITS + 2 * MGROUPS;
    // 64-bit exchanges in the transforms (half the shared-memory instructions of the planar form, and register pairs
    // arrive aligned for the packed f32x2 butterflies); their 4.6 KB of scratch per warp is paid for by keeping ONE mask
    // tile: the next tile is requested once all 16 warps have read the current one, a stage and a half ahead of its use
    static constexpr bool VEC = true;
    static_assert(TAIL > 0 && HEAD <= 16 && NB >= 1 && NB <= 2, "hop must be 128, 160 or 256");
    static size_t bytes() {
        return al16(sizeof(float2) * TW3N) + al16(sizeof(float) * UNITS * f3::Scr<VEC>::FLOATS) +
               al16(sizeof(float) * UNITS * 2 * TAIL * 32) + al16(sizeof(float) * UNITS * SEG) +
               al16(sizeof(float) * MASK_TILE) + al16(sizeof(uint64_t) * NBARS) + al16(sizeof(float) * 4 * NT);
    }
    static constexpr int TW3N = f3::TW1024_OFF;
};

#ifndef ADV_EXPLAIN4_MAXREG
// Register cap, measured on B200 (64 x 4 s clips; kernel alone / pooled step with the normaliser and the metric
// reduction): 128 -> 59.7 / 70.9 us; 120 (leaves room for one 128-thread normaliser CTA per SM, which then overlaps
// the next batch's explain) -> 63.9 / 72.5 us; 112 (one 256-thread normaliser CTA) -> 68.6 / 76.4 us: below 128 the
// transforms pay in spills and register-pair moves more than the overlap returns.
#define ADV_EXPLAIN4_MAXREG 128
#endif

template <int MODE, int HS>
__global__ void __maxnreg__(ADV_EXPLAIN4_MAXREG)
explain4_kernel(PlanDev P, int upc, int total_units, const float* __restrict__ wav, int64_t wav_stride,
                const float* __restrict__ mask, int Fm, int Tm, int drop, float* __restrict__ rel, float* __restrict__ irr,
                double* __restrict__ stats, int slots) {
    using C = E4Cfg<HS>;
    constexpr int UNITS = C::UNITS, NT = C::NT, F = C::F, MP = C::MP, HOP = C::HOP, USTEP = C::USTEP;
    constexpr int ROWS = C::ROWS, HEAD = C::HEAD, TAIL = C::TAIL, NB = C::NB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Carver cv{smem_raw};
    float2* tw_s = cv.take<float2>(C::TW3N);
    float* scratch = cv.take<float>(UNITS * f3::Scr<C::VEC>::FLOATS);
    float* tails = cv.take<float>(UNITS * 2 * TAIL * 32);
    float* seg_all = cv.take<float>(UNITS * C::SEG);
    float* mask_s = cv.take<float>(C::MASK_TILE);
    uint64_t* bars = cv.take<uint64_t>(C::NBARS);
    // per-thread statistics accumulators live in shared memory ([4][NT]): held in registers across the transforms they were
    // what the compiler spilled, and a local-memory reload misses the small L1 this kernel leaves (2 % of the warp samples
    // sat on it; explain5_kernel: 6 %)
    float* acc_s = cv.take<float>(4 * NT) + threadIdx.x;
    uint64_t* full = bars;                 // [16] tail of unit slot w written (1 arrival per pass)
    uint64_t* empty = bars + UNITS;        // [16] tail of unit slot w read by its NB consumers
    uint64_t* segbar = bars + 2 * UNITS;   // [16] waveform slice of warp w landed
    constexpr int MGW = C::MGW, MGROUPS = C::MGROUPS, MNT = 32 * MGW, MGC = 2 * MGW;
    uint64_t* mfull_all = bars + 3 * UNITS;      // [groups] the group's mask columns landed (32 MGW cp.async arrivals per pass)
    uint64_t* mempty_all = mfull_all + MGROUPS;  // [groups] ... read by the group's MGW warps

    const int tid = threadIdx.x, l = tid & 31;
    const int w = __shfl_sync(0xffffffffu, tid >> 5, 0);  // (tells the compiler the warp index is warp-uniform)

    // this CTA's run of the unit list: [g_begin, g_end), preceded by `halo` units transformed for their tails only
    const long G = gridDim.x;
    const int g_begin = (int)(((long)blockIdx.x * total_units) / G);
    const int g_end = (int)(((long)(blockIdx.x + 1) * total_units) / G);
    const int b0 = g_begin / upc, u0 = g_begin - b0 * upc;
    const int halo = u0 < NB ? u0 : NB;
    const int start = g_begin - halo;
    const int n_pass = (g_end - start + UNITS - 1) / UNITS;
    const int T_eff = drop ? (Tm < P.T ? Tm : P.T) : P.T;  // frames past the mask are dropped from both outputs
    const int f_lim = drop ? Fm : F;

    if (tid == 0) {
        for (int i = 0; i < UNITS; ++i) {
            mbar_init(full + i, 1);
            mbar_init(empty + i, NB);
            mbar_init(segbar + i, 1);
        }
        for (int i = 0; i < MGROUPS; ++i) {
            mbar_init(mfull_all + i, MNT);
            mbar_init(mempty_all + i, MGW);
        }
    }
    for (int i = tid; i < C::TW3N / 2; i += NT) cp_async16(tw_s + 2 * i, P.tw3 + 2 * i);
    cp_async_commit();
    pdl_launch_dependents();
    cp_async_wait_all();
    __syncthreads();  // tables staged, barriers initialised (the only CTA-wide barrier of the kernel)
    pdl_wait();

    // ---- statistics slots: zero this warp's slot of every clip the run touches; the first CTA of a clip also zeroes
    // the slots no CTA owns
    if (stats != nullptr && g_end > g_begin) {
        const int b_last = (g_end - 1) / upc;
        for (int b = b0; b <= b_last; ++b) {
            const long gb = (long)b * upc;
            const int c_first = (int)(((gb + 1) * G - 1) / total_units);
            const int c_last = (int)(((gb + upc) * G - 1) / total_units);
            double* row = stats + (size_t)b * slots * 4;
            if (l < 4) row[(((int)blockIdx.x - c_first) * UNITS + w) * 4 + l] = 0.0;
            if ((int)blockIdx.x == c_first)
                for (int i = (c_last - c_first + 1) * UNITS * 4 + tid; i < slots * 4; i += NT) row[i] = 0.0;
        }
    }

    float* my = scratch + w * f3::Scr<C::VEC>::FLOATS;
    float* seg = seg_all + w * C::SEG;
    float* tail_w = tails + w * (2 * TAIL * 32);
    const int q1 = l == 0 ? 32 : 64 - l;
    constexpr int seglen = HOP + 512;

    // ---- per-thread mask staging: column c of the tile <-> frame (c & 1) of unit slot c >> 1; rows f0 + 16 k
    //      (the columns are refilled per group of MGW adjacent warps: a group's 32 MGW threads copy its 2 MGW columns)
    const int mgrp = w / MGW, mtg = tid - mgrp * MNT;
    const int mc = mgrp * MGC + mtg % MGC, mf0 = mtg / MGC;
    uint64_t* mfull = mfull_all + mgrp;
    uint64_t* mempty = mempty_all + mgrp;
    UnitPos mpos{b0, u0};           // unit of this thread's mask column in the pass being requested
    {
        int g = start + (mc >> 1);  // may precede g_begin (halo units have mask columns too)
        mpos.b = g / upc;
        mpos.u = g - mpos.b * upc;
    }
    auto request_mask = [&](int pass) {
        const int g = start + pass * UNITS + (mc >> 1);
        const int t = 2 * mpos.u + (mc & 1);
        const bool col_ok = g < g_end && t < Tm;
        const float* mrow = mask + (size_t)mpos.b * Fm * Tm;
        const float* src = col_ok ? mrow + (size_t)mf0 * Tm + t : mask;
        const size_t step = col_ok ? (size_t)16 * Tm : 0;
        uint32_t dst = smem_u32(mask_s + mf0 * MP + mc);
        const int full_rows = col_ok ? (Fm - mf0 + 15) / 16 : 0;   // trips whose row exists in the mask
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int sz = k < full_rows ? 4 : 0;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + k * 16 * MP * 4), "l"(src), "r"(sz) : "memory");
            src += step;
        }
        if (mf0 == 0) {  // row 256
            const int sz = 16 < full_rows ? 4 : 0;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + 256 * MP * 4), "l"(src), "r"(sz) : "memory");
        }
        cp_async_arrive_noinc(mfull);
        mpos.advance(UNITS, upc);
    };

    // ---- this warp's unit, pass by pass
    UnitPos pos;
    {
        const int g = start + w;
        pos.b = g / upc;
        pos.u = g - pos.b * upc;
    }
    auto request_seg = [&](const UnitPos& q, bool live) -> int {
        // (a unit past the run still arms the barrier: zero-byte request)
        const int base = 2 * q.u * HOP - 256;
        if (live) return stage_segment_async<32>(seg, seglen, wav + (size_t)q.b * wav_stride, base, P.n_in, segbar + w, l);
        if (l == 0) mbar_expect_tx(segbar + w, 0);
        return 0;
    };

    request_mask(0);
    int shift = request_seg(pos, start + w < g_end);

#pragma unroll
    for (int i = 0; i < 4; ++i) acc_s[i * NT] = 0.f;   // (sum, sum sq) of rel, irr stored by this lane for clip acc_b
    int acc_b = -1;
    auto flush = [&]() {   // warp-uniform call sites
        double q[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) q[i] = warp_sum((double)acc_s[i * NT]);
        if (stats != nullptr && acc_b >= 0 && l == 0) {
            const int c_first = (int)((((long)acc_b * upc + 1) * G - 1) / total_units);
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
#if ADV_E4_EARLY_EXIT
        // (last pass only; nobody waits for anything such a warp would arrive on: its slot's consumers are idle as well, the
        //  mask tile is not refilled any more)
        if (!active) break;
#endif
        const bool is_out = active && g >= g_begin;
        const UnitPos cur = pos;
        const int cur_shift = shift;
        const int fa = 2 * cur.u;

        // -- 1. samples of frames a (real part) and b (imaginary part)
        float2 v[16], vi[16];
        __syncwarp();
        mbar_wait(segbar + w, p & 1);
        if (active) {
            const float* sa = seg + cur_shift + l;
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = make_float2(sa[32 * j], sa[HOP + 32 * j]);
            if (fa + 1 >= T_eff) {   // (warp-uniform, last unit of a clip) frames that do not exist / are dropped
                const bool ka = fa < T_eff;
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = make_float2(ka ? v[j].x : 0.f, 0.f);
            }
        } else {   // a unit past the end of the run (last pass only): transformed as silence, nothing stored.  No branch
                   // encloses the transforms: a thread-index-derived condition around a __syncwarp() costs a WARPSYNC /
                   // ENDCOLLECTIVE sequence per sync even when it is warp-uniform.
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = make_float2(0.f, 0.f);
        }
        pos.advance(UNITS, upc);

        {
            // -- 2. forward transform, both frames at once
            //    (the next pass's slice is requested from inside the transform, once every lane has consumed its samples)
            f3::fft_forward<C::VEC>(v, l, tw_s, my, [&] { if (p + 1 < n_pass) shift = request_seg(pos, g + UNITS < g_end); });
            float2 xa[9], xb[9];
            f3::split(v, l, xa, xb);
            if (f_lim < F) {   // (uniform) outside="drop" with a mask narrower than the spectrum
#pragma unroll
                for (int i = 0; i < 9; ++i) {
                    const int bin = i < 4 ? l + 64 * i : (i < 8 ? q1 + 64 * (i - 4) : 256);
                    if (bin >= f_lim) xa[i] = xb[i] = make_float2(0.f, 0.f);
                }
            }
            // -- 3. mask columns of the two frames, gains, the two inverse-transform inputs
            mbar_wait(mfull, p & 1);
            float ma[9], mb[9];
            {
                const float* p0 = mask_s + l * MP + 2 * w;
                const float* p1 = mask_s + q1 * MP + 2 * w;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    ma[i] = p0[64 * MP * i];
                    mb[i] = p0[64 * MP * i + 1];
                    ma[4 + i] = p1[64 * MP * i];
                    mb[4 + i] = p1[64 * MP * i + 1];
                }
                ma[8] = mask_s[256 * MP + 2 * w];
                mb[8] = mask_s[256 * MP + 2 * w + 1];
            }
            float2 yra[9], yia[9], yrb[9], yib[9];
            frame_gains<MODE>(xa, ma, yra, yia);
            frame_gains<MODE>(xb, mb, yrb, yib);
            f3::merge(v, l, yra, yrb);    // masked-in:  Z = Ya + i Yb
            f3::merge(vi, l, yia, yib);   // masked-out
        }
        __syncwarp();
        if (l == 0) mbar_arrive(mempty);   // (the gains above consumed the mask values: the loads have returned)

        // -- 5. inverse transforms; overlap-add of the unit's two frames in registers: strip row r (32 samples, lane l)
        //       = frame a row r (r < 16) + frame b row r - HS (r >= HS).  Rows < HEAD stay in registers, the tail rows go
        //       to shared memory for the next NB units.
        float head_r[HEAD], head_i[HEAD];
#if !ADV_E4_LATE_EMPTY
        if (p >= 1) mbar_wait(empty + w, (p - 1) & 1);   // the previous pass's tail of this slot has been consumed
#endif
        {
            f3::fft_inverse<C::VEC>(v, l, tw_s, my);
#if ADV_E4_LATE_EMPTY
            if (p >= 1) mbar_wait(empty + w, (p - 1) & 1);   // the previous pass's tail of this slot has been consumed
#endif
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                float o = r < 16 ? v[r < 16 ? r : 0].x : 0.0f;
                if (r >= HS) o += v[r >= HS ? r - HS : 0].y;
                if (r < HEAD) head_r[r < HEAD ? r : 0] = o;
                else tail_w[(r - HEAD) * 32 + l] = o;
            }
            f3::fft_inverse<C::VEC>(vi, l, tw_s, my);
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                float o = r < 16 ? vi[r < 16 ? r : 0].x : 0.0f;
                if (r >= HS) o += vi[r >= HS ? r - HS : 0].y;
                if (r < HEAD) head_i[r < HEAD ? r : 0] = o;
                else tail_w[TAIL * 32 + (r - HEAD) * 32 + l] = o;
            }
        }
        __syncwarp();
        if (l == 0) mbar_arrive(full + w);
        // -- this thread's share of the next pass's mask tile, once every warp has read the current one (the slowest warp
        //    finished its gains about two transforms ago; the copy has the rest of this pass and the next forward
        //    transform to land)
        if (p + 1 < n_pass) {
            mbar_wait(mempty, p & 1);
            request_mask(p + 1);
        }

        // reciprocal envelope of the head rows (requested before the waits below)
        const int s_base = USTEP * cur.u - 256 + l;   // output sample of head row 0
        float env[HEAD];
        if (is_out) {
#pragma unroll
            for (int r = 0; r < HEAD; ++r) {
                const int sidx = s_base + 32 * r;
                env[r] = (sidx >= 0 && sidx < P.n_out) ? __ldg(P.inv_env + sidx) : 0.0f;
            }
        }

        // -- 6. tails of the previous NB units (unit slot w - k; the slots before slot 0 belong to the previous pass).
        //       Farthest neighbour first: the arrival on the nearest slot's `empty` barrier is this warp's last access to
        //       any tail of the pass.  The waits are unconditional so that barrier phases advance in lock-step.
#pragma unroll
        for (int k = NB; k >= 1; --k) {
            const int slot = (w - k) & (UNITS - 1);
            const int pp = w >= k ? p : p - 1;
            {
                if (pp >= 0) mbar_wait(full + slot, pp & 1);
                if (is_out && cur.u >= k) {
                    const float* tn = tails + slot * (2 * TAIL * 32) + l;
                    constexpr int r0 = 0;
#pragma unroll
                    for (int r = (k - 1) * HEAD; r < k * HEAD && r < TAIL; ++r) {
                        head_r[r - (k - 1) * HEAD + r0] += tn[r * 32];
                        head_i[r - (k - 1) * HEAD + r0] += tn[TAIL * 32 + r * 32];
                    }
                }
                __syncwarp();
                if (l == 0 && pp >= 0) mbar_arrive(empty + slot);
            }
        }

        // -- 7. scale, statistics, store (rows of 32 consecutive samples: 128-byte stores)
        if (is_out) {
            if (cur.b != acc_b) {   // (warp-uniform)
                if (acc_b >= 0) flush();
                acc_b = cur.b;
            }
            float* rrow = rel + (size_t)cur.b * P.n_out;
            float* irow = irr + (size_t)cur.b * P.n_out;
            float acc[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = acc_s[i * NT];
#pragma unroll
            for (int r = 0; r < HEAD; ++r) {
                const int sidx = s_base + 32 * r;
                const float a = head_r[r] * env[r], c = head_i[r] * env[r];   // (env = 0 outside [0, n_out))
                acc[0] += a;
                acc[1] = fmaf(a, a, acc[1]);
                acc[2] += c;
                acc[3] = fmaf(c, c, acc[3]);
                if (sidx >= 0 && sidx < P.n_out) {
                    rrow[sidx] = a;
                    irow[sidx] = c;
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) acc_s[i * NT] = acc[i];
        }
    }
    if (acc_b >= 0) flush();
}

// ------------------------------------------------------------------------------------------------------------------
// Streaming iSTFT (torch.istft, audioprocessor.py:123-129) - the same pipeline of warps without the forward transform
// and the mask: a unit's two spectrum rows come straight from global memory into the lanes that own their bins
// (32 consecutive bins per load instruction), ONE inverse transform carries frames a and b as real / imaginary part,
// overlap-add in registers, tails through shared memory, heads stored from registers.  64 registers per thread and
// the next unit's rows are requested (into registers) as soon as the current ones are merged, a whole pass ahead of use.
// ------------------------------------------------------------------------------------------------------------------
#ifndef ADV_ISTFT4_UNITS
#define ADV_ISTFT4_UNITS 8   // warps (= units per pass) of a CTA
#endif
#ifndef ADV_ISTFT4_CTAS
#define ADV_ISTFT4_CTAS 2    // resident CTAs per SM the kernel is compiled for: 2 x 256 threads -> up to 128 registers per thread
#endif                       // (the next unit's 18 spectrum values per lane are held in registers across the transform).
                             // Measured (64 x 4 s clips, rows loaded at the start of a pass): 2 x 512 threads at 64 registers
                             // spills, 29.9 us; 1 x 512 at 128: 21.1 us; 3 x 256 at 80: 20.7 us; 2 x 256 at 120: 20.3 us
constexpr int kI4Units = ADV_ISTFT4_UNITS, kI4Threads = 32 * kI4Units;

#ifndef ADV_I4_TAIL_BUFS
#define ADV_I4_TAIL_BUFS 2
#endif
template <int HS>
struct I4Cfg {
    using E = E4Cfg<HS>;
    // TB tail buffers per unit slot (2: a warp writes pass p's tail while the next warps may still be reading pass p - 1's -
    // shared memory is not what limits this kernel's occupancy)
    static constexpr int UNITS = kI4Units, TB = ADV_I4_TAIL_BUFS, NBARS = 2 * TB * UNITS;
    static size_t bytes() {
        return al16(sizeof(float2) * E::TW3N) + al16(sizeof(float) * UNITS * f3::Scr<true>::FLOATS) +
               al16(sizeof(float) * TB * UNITS * E::TAIL * 32) + al16(sizeof(uint64_t) * NBARS);
    }
};

template <int HS, bool CONTIG>
__global__ void __launch_bounds__(kI4Threads, ADV_ISTFT4_CTAS)
istft4_kernel(PlanDev P, int upc, int total_units, const float2* __restrict__ X, int64_t sb, int64_t st, int64_t sf,
              float* __restrict__ out, double* __restrict__ stats, int slots) {
    using C = E4Cfg<HS>;
    constexpr int UNITS = kI4Units, NT = kI4Threads, HOP = C::HOP, USTEP = C::USTEP;
    constexpr int ROWS = C::ROWS, HEAD = C::HEAD, TAIL = C::TAIL, NB = C::NB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Carver cv{smem_raw};
    float2* tw_s = cv.take<float2>(C::TW3N);
    float* scratch = cv.take<float>(UNITS * f3::Scr<true>::FLOATS);
    constexpr int TB = I4Cfg<HS>::TB;
    float* tails = cv.take<float>(TB * UNITS * TAIL * 32);            // [TB][UNITS][TAIL][32]
    uint64_t* bars = cv.take<uint64_t>(I4Cfg<HS>::NBARS);
    uint64_t* full = bars;                    // [TB][UNITS] tail of unit slot w written (pass p uses buffer p % TB)
    uint64_t* empty = bars + TB * UNITS;      // [TB][UNITS] ... read by its NB consumers

    const int tid = threadIdx.x, l = tid & 31;
    const int w = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const long G = gridDim.x;
    const int g_begin = (int)(((long)blockIdx.x * total_units) / G);
    const int g_end = (int)(((long)(blockIdx.x + 1) * total_units) / G);
    const int b0 = g_begin / upc, u0 = g_begin - b0 * upc;
    const int halo = u0 < NB ? u0 : NB;
    const int start = g_begin - halo;
    const int n_pass = (g_end - start + UNITS - 1) / UNITS;

    if (tid == 0)
        for (int i = 0; i < TB * UNITS; ++i) {
            mbar_init(full + i, 1);
            mbar_init(empty + i, NB);
        }
    for (int i = tid; i < C::TW3N / 2; i += NT) cp_async16(tw_s + 2 * i, P.tw3 + 2 * i);
    cp_async_commit();
    pdl_launch_dependents();
    cp_async_wait_all();
    __syncthreads();
    pdl_wait();

    if (stats != nullptr && g_end > g_begin) {
        const int b_last = (g_end - 1) / upc;
        for (int b = b0; b <= b_last; ++b) {
            const long gb = (long)b * upc;
            const int c_first = (int)(((gb + 1) * G - 1) / total_units);
            const int c_last = (int)(((gb + upc) * G - 1) / total_units);
            double* row = stats + (size_t)b * slots * 2;
            if (l < 2) row[(((int)blockIdx.x - c_first) * UNITS + w) * 2 + l] = 0.0;
            if ((int)blockIdx.x == c_first)
                for (int i = (c_last - c_first + 1) * UNITS * 2 + tid; i < slots * 2; i += NT) row[i] = 0.0;
        }
    }

    float* my = scratch + w * f3::Scr<true>::FLOATS;
    const int q1 = l == 0 ? 32 : 64 - l;
    const int64_t sfe = CONTIG ? 1 : sf;

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
            const int c_first = (int)((((long)acc_b * upc + 1) * G - 1) / total_units);
            double* row = stats + ((size_t)acc_b * slots + ((int)blockIdx.x - c_first) * UNITS + w) * 2;
            row[0] = q0;
            row[1] = q1s;
        }
        acc[0] = acc[1] = 0.f;
    };
    const float2 zero2 = make_float2(0.f, 0.f);
    // spectrum rows of a unit -> the lanes that own their bins (frames past the last one, units past the run: zeros)
    float2 xa[9], xb[9];
    auto load_rows = [&](const UnitPos& q, bool live) {
        const int fa = 2 * q.u;
        const bool va = live && fa < P.T, vb = live && fa + 1 < P.T;
        const float2* pa = X + (size_t)q.b * sb + (size_t)fa * st;
        const float2* pb = pa + st;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            xa[i] = va ? __ldg(pa + (l + 64 * i) * sfe) : zero2;
            xa[4 + i] = va ? __ldg(pa + (q1 + 64 * i) * sfe) : zero2;
            xb[i] = vb ? __ldg(pb + (l + 64 * i) * sfe) : zero2;
            xb[4 + i] = vb ? __ldg(pb + (q1 + 64 * i) * sfe) : zero2;
        }
        xa[8] = (l == 0 && va) ? __ldg(pa + 256 * sfe) : zero2;
        xb[8] = (l == 0 && vb) ? __ldg(pb + 256 * sfe) : zero2;
    };
    load_rows(pos, start + w < g_end);

    for (int p = 0; p < n_pass; ++p) {
        const int g = start + p * UNITS + w;
        const bool active = g < g_end;
#if ADV_E4_EARLY_EXIT
        if (!active) break;   // (last pass only; see explain4_kernel)
#endif
        const bool is_out = active && g >= g_begin;
        const UnitPos cur = pos;
        pos.advance(UNITS, upc);

        // -- 1. the unit's two spectrum rows were requested a pass ago (registers xa / xb); build the transform input and
        //       request the next pass's rows, which then travel while this pass computes
        float2 v[16];
        f3::merge(v, l, xa, xb);
        load_rows(pos, p + 1 < n_pass && g + UNITS < g_end);
        // -- 2. inverse transform, overlap-add of the two frames in registers, tail rows to shared memory
        f3::fft_inverse<true>(v, l, tw_s, my);
        float head[HEAD];
        // buffer p % TB of this slot was last written in pass p - TB: its phase index on either barrier is p / TB
        const int tb = p % TB;
        float* tail_w = tails + (tb * UNITS + w) * (TAIL * 32);
        if (p >= TB) mbar_wait(empty + tb * UNITS + w, (p / TB - 1) & 1);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            float o = r < 16 ? v[r < 16 ? r : 0].x : 0.0f;
            if (r >= HS) o += v[r >= HS ? r - HS : 0].y;
            if (r < HEAD) head[r < HEAD ? r : 0] = o;
            else tail_w[(r - HEAD) * 32 + l] = o;
        }
        __syncwarp();
        if (l == 0) mbar_arrive(full + tb * UNITS + w);

        const int s_base = USTEP * cur.u - 256 + l;
        float env[HEAD];
        if (is_out) {
#pragma unroll
            for (int r = 0; r < HEAD; ++r) {
                const int sidx = s_base + 32 * r;
                env[r] = (sidx >= 0 && sidx < P.n_out) ? __ldg(P.inv_env + sidx) : 0.0f;
            }
        }
        // -- 3. tails of the previous NB units (see explain4_kernel for the protocol)
#pragma unroll
        for (int k = NB; k >= 1; --k) {
            const int slot = (w - k) & (UNITS - 1);
            const int pp = w >= k ? p : p - 1;
            const int pb = pp >= 0 ? pp % TB : 0;
            if (pp >= 0) mbar_wait(full + pb * UNITS + slot, (pp / TB) & 1);
            if (is_out && cur.u >= k) {
                const float* tn = tails + (pb * UNITS + slot) * (TAIL * 32) + l;
#pragma unroll
                for (int r = (k - 1) * HEAD; r < k * HEAD && r < TAIL; ++r) head[r - (k - 1) * HEAD] += tn[r * 32];
            }
            __syncwarp();
            if (l == 0 && pp >= 0) mbar_arrive(empty + pb * UNITS + slot);
        }
        // -- 4. scale, statistics, store
        if (is_out) {
            if (cur.b != acc_b) {
                if (acc_b >= 0) flush();
                acc_b = cur.b;
            }
            float* orow = out + (size_t)cur.b * P.n_out;
#pragma unroll
            for (int r = 0; r < HEAD; ++r) {
                const int sidx = s_base + 32 * r;
                const float a = head[r] * env[r];
                acc[0] += a;
                acc[1] = fmaf(a, a, acc[1]);
                if (sidx >= 0 && sidx < P.n_out) orow[sidx] = a;
            }
        }
    }
    if (acc_b >= 0) flush();
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
static bool gen4_enabled() {
    static const char* e = ADV_AB_ENV("ADV_GEN4");   // A/B switch: ADV_GEN4=0 routes the call to the generation-3 kernel
    static const bool on = !(e && e[0] == '0');
    return on;
}

static int sm_count4() {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    static int cache[64] = {0};
    if (dev >= 0 && dev < 64 && cache[dev] > 0) return cache[dev];
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    if (dev >= 0 && dev < 64) cache[dev] = n;
    return n;
}

struct Run4 { int hs, upc, grid, slots; long total; };
// geometry of a generation-4 launch (hs = 0: outside the kernel's domain)
static Run4 plan_run4(const adv_plan* p, int batch) {
    Run4 r = {0, 0, 0, 0, 0};
    const PlanDev& d = p->d;
    if (!gen4_enabled() || d.n_fft != 512 || !d.rect_full || d.n_in <= 0 || d.n_out <= 0 || batch <= 0) return r;
    if (d.hop != 128 && d.hop != 160 && d.hop != 256) return r;
    r.upc = (d.n_out + 256 + 2 * d.hop - 1) / (2 * d.hop);   // units whose heads cover every output sample
    r.total = (long)r.upc * batch;
    if (r.total > 0x3fffffffL) return r;
    const long by_work = r.total / 8 > 0 ? r.total / 8 : 1;   // at least 8 units per CTA (2 of a run are halo)
    r.grid = (int)(by_work < sm_count4() ? by_work : sm_count4());
    // CTAs that can touch one clip: its first and last unit are upc - 1 apart in a list cut into `grid` equal runs
    const long span = ((long)(r.upc - 1) * r.grid + r.total - 1) / r.total;
    r.slots = 16 * (int)(span + 1);
    r.hs = d.hop / 32;
    return r;
}

int explain4_slots(const adv_plan* p, int batch) {
    const Run4 r = plan_run4(p, batch);
    return r.hs ? r.slots : 0;
}

template <int MODE, int HS>
static int launch_explain4_t(const adv_plan* p, const Run4& r, const float* wav, int64_t wav_stride, const float* mask,
                             int Fm, int Tm, int drop, float* rel, float* irr, double* stats, cudaStream_t s) {
    const size_t smem = E4Cfg<HS>::bytes();
    auto kernel = explain4_kernel<MODE, HS>;
    int rc = set_smem(kernel, smem);
    if (rc != ADV_OK) return rc;
    ADV_CUDA_CHECK(launch_pdl(kernel, r.grid, 512, smem, s, p->d, r.upc, (int)r.total, wav, wav_stride, mask, Fm, Tm, drop,
                              rel, irr, stats, r.slots));
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

int launch_explain4(const adv_plan* p, const float* wav, int64_t wav_stride, const float* mask, int Fm, int Tm,
                    int mode_flags, int batch, float* rel, float* irr, double* stats, cudaStream_t s) {
    const Run4 r = plan_run4(p, batch);
    if (!r.hs || wav == nullptr) return ADV_ERR_UNSUPPORTED;
    const int mode = mode_flags & 0xff, drop = (mode_flags & ADV_MASK_DROP_OUTSIDE) ? 1 : 0;
#define ADV_E4(HS)                                                                                                        \
    return mode == ADV_MASK_LOG1P                                                                                         \
               ? launch_explain4_t<ADV_MASK_LOG1P, HS>(p, r, wav, wav_stride, mask, Fm, Tm, drop, rel, irr, stats, s)     \
               : launch_explain4_t<ADV_MASK_LINEAR, HS>(p, r, wav, wav_stride, mask, Fm, Tm, drop, rel, irr, stats, s)
    if (r.hs == 4) ADV_E4(4);
    if (r.hs == 5) ADV_E4(5);
    ADV_E4(8);
#undef ADV_E4
}

// ---- streaming iSTFT
struct RunI4 { int hs, upc, grid, slots; long total; };
static RunI4 plan_run_i4(const adv_plan* p, int batch) {
    RunI4 r = {0, 0, 0, 0, 0};
    const PlanDev& d = p->d;
    if (!gen4_enabled() || d.n_fft != 512 || !d.rect_full || d.n_out <= 0 || batch <= 0) return r;
    if (d.hop != 128 && d.hop != 160 && d.hop != 256) return r;
    r.upc = (d.n_out + 256 + 2 * d.hop - 1) / (2 * d.hop);
    r.total = (long)r.upc * batch;
    if (r.total > 0x3fffffffL) return r;
    const long by_work = r.total / 8 > 0 ? r.total / 8 : 1;
    const long slots_hw = (long)ADV_ISTFT4_CTAS * sm_count4();
    r.grid = (int)(by_work < slots_hw ? by_work : slots_hw);
    const long span = ((long)(r.upc - 1) * r.grid + r.total - 1) / r.total;
    r.slots = kI4Units * (int)(span + 1);
    r.hs = d.hop / 32;
    return r;
}

int istft4_slots(const adv_plan* p, int batch) {
    const RunI4 r = plan_run_i4(p, batch);
    return r.hs ? r.slots : 0;
}

template <int HS, bool CONTIG>
static int launch_istft4_t(const adv_plan* p, const RunI4& r, const float2* X, int64_t sb, int64_t st, int64_t sf,
                           float* out, double* stats, cudaStream_t s) {
    const size_t smem = I4Cfg<HS>::bytes();
    auto kernel = istft4_kernel<HS, CONTIG>;
    int rc = set_smem(kernel, smem);
    if (rc != ADV_OK) return rc;
    ADV_CUDA_CHECK(launch_pdl(kernel, r.grid, kI4Threads, smem, s, p->d, r.upc, (int)r.total, X, sb, st, sf, out, stats, r.slots));
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

int launch_istft4(const adv_plan* p, const float2* X, int64_t sb, int64_t st, int64_t sf, int batch, float* out,
                  double* stats, cudaStream_t s) {
    const RunI4 r = plan_run_i4(p, batch);
    if (!r.hs) return ADV_ERR_UNSUPPORTED;
#define ADV_I4(HS)                                                                        \
    return sf == 1 ? launch_istft4_t<HS, true>(p, r, X, sb, st, sf, out, stats, s)        \
                   : launch_istft4_t<HS, false>(p, r, X, sb, st, sf, out, stats, s)
    if (r.hs == 4) ADV_I4(4);
    if (r.hs == 5) ADV_I4(5);
    ADV_I4(8);
#undef ADV_I4
}

}  // namespace adv
