// Generation-3 warp FFT: 512 complex points on 32 lanes x 16 registers as 8 x 8 x 8, NO shuffles.
//
// Why (ncu + SASS accounting of the generation-2 "wide" unit, scripts/sass_lines.py): a 512-point transform cost
// 700 - 820 issued instructions against ~430 of arithmetic.  The excess was the 16 x 32 factorisation itself: rows
// of 32 points split over a lane pair (32 SHFL for the radix-2 across the pair, a predicated twiddle pass that every
// lane issues), mirror bins k / 512-k living in different lanes (16 more SHFL per split / merge plus selects), and -
// because the kernels' persistent loops are not provably warp-uniform - a WARPSYNC + ENDCOLLECTIVE pair around
// every shuffle.  Here:
//   n = n0 + 8 n1 + 64 n2,  k = k2 + 8 k1 + 64 k0:
//   A: DFT-8 over n2 (registers)          x W64^(n1 k2)         -> transpose through shared memory
//   B: DFT-8 over n1 (registers)          x W512^(n0 (k2+8k1))  -> transpose through shared memory
//   C: DFT-8 over n0 (registers)          -> Z[q + 64 k0], q = k2 + 8 k1
//   * time side:      lane l holds z[l + 32 j], j = 0..15  (register 2 n2 + c <-> n = l + 32 c + 64 n2): the layout
//                     the generation-2 kernels already stage, window and overlap-add in;
//   * frequency side: lane l holds columns q = l and q = 64 - l (lane 0: 0 and 32), eight bins each.  The mirror
//                     of bin l + 64 k0 is (64 - l) + 64 (7 - k0): SAME LANE.  Splitting two packed real frames,
//                     merging two Hermitian spectra and the half-size real-FFT post-processing (n_fft 1024 frames as
//                     512-point complex transforms) are register-local; only lane 0 needs selects.
//   * every store / load instruction of a warp covers 32 consecutive bins (256 contiguous bytes of a spectrum row).
// Transposes are 64-bit (VEC = true: 16 + 16 shared-memory instructions per exchange, 4.6 KB of scratch per warp) or
// planar (VEC = false: 32 + 32 instructions, 2.3 KB) - all access patterns conflict-free (pitches 72 / 66 / 68).
// Twiddles come from a 4.7 KB shared table (broadcast or conflict-free 64-bit loads).
//
// Everything is __host__ __device__ and split into per-lane steps so tests/host_emul.cu runs the same code lane by
// lane on the CPU against a float64 DFT.
#pragma once
#include <type_traits>
#include "fft_core.cuh"

namespace adv {
namespace f3 {

constexpr int P1 = 72;           // pitch of exchange 1, [k2][m], m = n0 + 8 n1
constexpr int P2V = 66;          // pitch of exchange 2, [n0][q], 64-bit accesses
constexpr int P2S = 68;          //                              planar accesses
constexpr int SCR_FLOATS_VEC = 2 * 8 * P1;   // float2 [8][72]
constexpr int SCR_FLOATS_PLANAR = 8 * P1;    // float  [8][72] (one plane at a time)
template <bool VEC> struct Scr { static constexpr int FLOATS = VEC ? SCR_FLOATS_VEC : SCR_FLOATS_PLANAR; };

// twiddle table layout (float2 units), built on the host by build_tables()
constexpr int TW512_PITCH = 66;                       // [n0][q]: exp(-2 pi i n0 q / 512)
constexpr int TW64_OFF = 8 * TW512_PITCH;             // [n1][k2]: exp(-2 pi i n1 k2 / 64)
constexpr int TW1024_OFF = TW64_OFF + 64;             // [k], k = 0..256: exp(-2 pi i k / 1024)
constexpr int TW_TOTAL = TW1024_OFF + 264;            // (257 padded to a multiple of 8)

inline void build_tables(float2* t) {
    const double pi = 3.14159265358979323846;
    for (int i = 0; i < TW_TOTAL; ++i) t[i] = make_float2(1.0f, 0.0f);
    for (int n0 = 0; n0 < 8; ++n0)
        for (int q = 0; q < 64; ++q) {
            const double a = -2.0 * pi * (double)(n0 * q) / 512.0;
            t[n0 * TW512_PITCH + q] = make_float2((float)cos(a), (float)sin(a));
        }
    for (int n1 = 0; n1 < 8; ++n1)
        for (int k2 = 0; k2 < 8; ++k2) {
            const double a = -2.0 * pi * (double)(n1 * k2) / 64.0;
            t[TW64_OFF + n1 * 8 + k2] = make_float2((float)cos(a), (float)sin(a));
        }
    for (int k = 0; k <= 256; ++k) {
        const double a = -2.0 * pi * (double)k / 1024.0;
        t[TW1024_OFF + k] = make_float2((float)cos(a), (float)sin(a));
    }
}

ADV_HD int q0_of(int l) { return l; }
ADV_HD int q1_of(int l) { return l == 0 ? 32 : 64 - l; }
constexpr int SLOTS = 9;  // one-sided bins of the 512-point transform per lane (8, + bin 256 on lane 0)
// one-sided bin (k <= 256) of lane l, slot i; -1 when empty
ADV_HD int bin_of(int l, int i) {
    if (i < 4) return l + 64 * i;
    if (i < 8) return q1_of(l) + 64 * (i - 4);
    return l == 0 ? 256 : -1;
}

// ---------------------------------------------------------------------------------------------------------------
// register stages
// ---------------------------------------------------------------------------------------------------------------
// forward A: v[2 n2 + c] -> v[8 c + k2] = (sum_n2 v W8^(n2 k2)) W64^(n1 k2),  n1 = (l >> 3) + 4 c
ADV_HD void fwd_a(float2* v, int l, const float2* tw) {
    float2 t0[8], t1[8];
    FFTReg<8, -1>::template run<2>(v, t0);
    FFTReg<8, -1>::template run<2>(v + 1, t1);
    const float2* w0 = tw + TW64_OFF + (l >> 3) * 8;
    const float2* w1 = w0 + 32;
    v[0] = t0[0];
    v[8] = t1[0];
#pragma unroll
    for (int k2 = 1; k2 < 8; ++k2) {
        v[k2] = cmul(t0[k2], w0[k2]);
        v[8 + k2] = cmul(t1[k2], w1[k2]);
    }
}
// forward B: v[8 c' + n1] -> v[8 c' + k1] = (sum_n1 v W8^(n1 k1)) W512^(n0 (k2 + 8 k1)),  n0 = l & 7, k2 = (l >> 3) + 4 c'
ADV_HD void fwd_b(float2* v, int l, const float2* tw) {
    float2 t0[8], t1[8];
    FFTReg<8, -1>::template run<1>(v, t0);
    FFTReg<8, -1>::template run<1>(v + 8, t1);
    const float2* w = tw + (l & 7) * TW512_PITCH + (l >> 3);
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) {
        v[k1] = cmul(t0[k1], w[8 * k1]);
        v[8 + k1] = cmul(t1[k1], w[8 * k1 + 4]);
    }
}
// forward C: v[8 g + n0] -> v[8 g + k0] = Z[q_g + 64 k0]
ADV_HD void fwd_c(float2* v) {
    fft_inplace<8, -1>(v);
    fft_inplace<8, -1>(v + 8);
}
// inverse C: v[8 g + k0] -> v[8 g + n0] = (sum_k0 v W8^(-n0 k0)) conj W512^(n0 q_g)
ADV_HD void inv_c(float2* v, int l, const float2* tw) {
    float2 t0[8], t1[8];
    FFTReg<8, +1>::template run<1>(v, t0);
    FFTReg<8, +1>::template run<1>(v + 8, t1);
    const float2* w0 = tw + q0_of(l);
    const float2* w1 = tw + q1_of(l);
    v[0] = t0[0];
    v[8] = t1[0];
#pragma unroll
    for (int n0 = 1; n0 < 8; ++n0) {
        v[n0] = cmulc(t0[n0], w0[n0 * TW512_PITCH]);
        v[8 + n0] = cmulc(t1[n0], w1[n0 * TW512_PITCH]);
    }
}
// inverse B: v[8 c' + k1] -> v[8 c' + n1] = (sum_k1 v W8^(-n1 k1)) conj W64^(n1 k2),  k2 = (l >> 3) + 4 c'
ADV_HD void inv_b(float2* v, int l, const float2* tw) {
    float2 t0[8], t1[8];
    FFTReg<8, +1>::template run<1>(v, t0);
    FFTReg<8, +1>::template run<1>(v + 8, t1);
    const float2* w = tw + TW64_OFF + (l >> 3);
    v[0] = t0[0];
    v[8] = t1[0];
#pragma unroll
    for (int n1 = 1; n1 < 8; ++n1) {
        v[n1] = cmulc(t0[n1], w[8 * n1]);
        v[8 + n1] = cmulc(t1[n1], w[8 * n1 + 4]);
    }
}
// inverse A: v[8 c + k2] -> v[2 n2 + c] = sum_k2 v W8^(-n2 k2)
ADV_HD void inv_a(float2* v) {
    float2 t0[8], t1[8];
    FFTReg<8, +1>::template run<1>(v, t0);
    FFTReg<8, +1>::template run<1>(v + 8, t1);
#pragma unroll
    for (int n2 = 0; n2 < 8; ++n2) {
        v[2 * n2] = t0[n2];
        v[2 * n2 + 1] = t1[n2];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// exchanges through shared memory.  X1: [k2][m] (m = n0 + 8 n1 = l + 32 c on the column side), X2: [n0][q].
// "cols" = the side indexed by the lane directly (consecutive lanes, consecutive addresses); "rows" = the side where
// a lane walks a row with stride 8 (X1) / 8 (X2, over k1).  VEC: float2 elements; planar: one float plane per pass.
// ---------------------------------------------------------------------------------------------------------------
template <bool VEC> struct X {
    static constexpr int P2 = VEC ? P2V : P2S;
    using T = typename std::conditional<VEC, float2, float>::type;
    static ADV_HD T get(const float2& a, bool imag) {
        if constexpr (VEC) return a; else return imag ? a.y : a.x;
    }
    static ADV_HD void put(float2& a, T x, bool imag) {
        if constexpr (VEC) a = x; else { if (imag) a.y = x; else a.x = x; }
    }
    // X1 column side: register 8 c + k2  <->  [k2][l + 32 c]
    static ADV_HD void x1_store_cols(const float2* v, int l, float* scr, bool imag) {
        T* s = reinterpret_cast<T*>(scr) + l;
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) {
            s[k2 * P1] = get(v[k2], imag);
            s[k2 * P1 + 32] = get(v[8 + k2], imag);
        }
    }
    static ADV_HD void x1_load_cols(float2* v, int l, const float* scr, bool imag) {
        const T* s = reinterpret_cast<const T*>(scr) + l;
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) {
            put(v[k2], s[k2 * P1], imag);
            put(v[8 + k2], s[k2 * P1 + 32], imag);
        }
    }
    // X1 row side: register 8 c' + n1  <->  [(l >> 3) + 4 c'][(l & 7) + 8 n1]
    static ADV_HD void x1_load_rows(float2* v, int l, const float* scr, bool imag) {
        const T* s = reinterpret_cast<const T*>(scr) + (l >> 3) * P1 + (l & 7);
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
            put(v[n1], s[8 * n1], imag);
            put(v[8 + n1], s[4 * P1 + 8 * n1], imag);
        }
    }
    static ADV_HD void x1_store_rows(const float2* v, int l, float* scr, bool imag) {
        T* s = reinterpret_cast<T*>(scr) + (l >> 3) * P1 + (l & 7);
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
            s[8 * n1] = get(v[n1], imag);
            s[4 * P1 + 8 * n1] = get(v[8 + n1], imag);
        }
    }
    // X2 row side: register 8 c' + k1  <->  [l & 7][(l >> 3) + 4 c' + 8 k1]
    static ADV_HD void x2_store_rows(const float2* v, int l, float* scr, bool imag) {
        T* s = reinterpret_cast<T*>(scr) + (l & 7) * P2 + (l >> 3);
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) {
            s[8 * k1] = get(v[k1], imag);
            s[8 * k1 + 4] = get(v[8 + k1], imag);
        }
    }
    static ADV_HD void x2_load_rows(float2* v, int l, const float* scr, bool imag) {
        const T* s = reinterpret_cast<const T*>(scr) + (l & 7) * P2 + (l >> 3);
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) {
            put(v[k1], s[8 * k1], imag);
            put(v[8 + k1], s[8 * k1 + 4], imag);
        }
    }
    // X2 column side: register 8 g + n0  <->  [n0][q_g]
    static ADV_HD void x2_load_cols(float2* v, int l, const float* scr, bool imag) {
        const T* s0 = reinterpret_cast<const T*>(scr) + q0_of(l);
        const T* s1 = reinterpret_cast<const T*>(scr) + q1_of(l);
#pragma unroll
        for (int n0 = 0; n0 < 8; ++n0) {
            put(v[n0], s0[n0 * P2], imag);
            put(v[8 + n0], s1[n0 * P2], imag);
        }
    }
    static ADV_HD void x2_store_cols(const float2* v, int l, float* scr, bool imag) {
        T* s0 = reinterpret_cast<T*>(scr) + q0_of(l);
        T* s1 = reinterpret_cast<T*>(scr) + q1_of(l);
#pragma unroll
        for (int n0 = 0; n0 < 8; ++n0) {
            s0[n0 * P2] = get(v[n0], imag);
            s1[n0 * P2] = get(v[8 + n0], imag);
        }
    }
};

#ifdef __CUDACC__
// Device-side composition.  `tw` = the shared-memory copy of the plan's generation-3 table, `scr` = the warp's private
// scratch (Scr<VEC>::FLOATS floats).  Both begin with a __syncwarp so that a previous exchange has been fully read.
// `after_inputs_consumed()` runs once every lane of the warp has USED all 16 of its input values (stage A reads them
// all, the warp sync joins the lanes): the place to let an asynchronous copy overwrite the buffer the inputs were
// loaded from.  A sync alone does not do that - a shared-memory load may still be queued when the next instruction
// issues - and a bulk copy landing before a queued load was seen on B200 as one stale 32-sample row per ~4000 units.
template <bool VEC, class Hook>
__device__ __forceinline__ void fft_forward(float2* v, int l, const float2* tw, float* scr, Hook&& after_inputs_consumed) {
    using E = X<VEC>;
    fwd_a(v, l, tw);
    __syncwarp();
    after_inputs_consumed();
    E::x1_store_cols(v, l, scr, false);
    __syncwarp();
    E::x1_load_rows(v, l, scr, false);
    if constexpr (!VEC) {
        __syncwarp();
        E::x1_store_cols(v, l, scr, true);
        __syncwarp();
        E::x1_load_rows(v, l, scr, true);
    }
    fwd_b(v, l, tw);
    __syncwarp();
    E::x2_store_rows(v, l, scr, false);
    __syncwarp();
    E::x2_load_cols(v, l, scr, false);
    if constexpr (!VEC) {
        __syncwarp();
        E::x2_store_rows(v, l, scr, true);
        __syncwarp();
        E::x2_load_cols(v, l, scr, true);
    }
    fwd_c(v);
}
template <bool VEC>
__device__ __forceinline__ void fft_forward(float2* v, int l, const float2* tw, float* scr) {
    fft_forward<VEC>(v, l, tw, scr, [] {});
}
template <bool VEC>
__device__ __forceinline__ void fft_inverse(float2* v, int l, const float2* tw, float* scr) {
    using E = X<VEC>;
    inv_c(v, l, tw);
    __syncwarp();
    E::x2_store_cols(v, l, scr, false);
    __syncwarp();
    E::x2_load_rows(v, l, scr, false);
    if constexpr (!VEC) {
        __syncwarp();
        E::x2_store_cols(v, l, scr, true);
        __syncwarp();
        E::x2_load_rows(v, l, scr, true);
    }
    inv_b(v, l, tw);
    __syncwarp();
    E::x1_store_rows(v, l, scr, false);
    __syncwarp();
    E::x1_load_cols(v, l, scr, false);
    if constexpr (!VEC) {
        __syncwarp();
        E::x1_store_rows(v, l, scr, true);
        __syncwarp();
        E::x1_load_cols(v, l, scr, true);
    }
    inv_a(v);
}
#endif

// ---------------------------------------------------------------------------------------------------------------
// Two real frames in one transform (n_fft 512): Z = FFT(xa + i xb).  Register-local: slot i pairs register a(i)
// with its mirror b(i); only lane 0 (columns 0 and 32, each its own mirror) selects different registers.
//   slots 0..3: bin l + 64 i        a = v[i]      b = v[15 - i]   (lane 0: v[(8 - i) & 7])
//   slots 4..7: bin q1 + 64 (i-4)   a = v[4 + i]  b = v[11 - i]   (lane 0: v[19 - i])
//   slot  8   : bin 256 (lane 0)    a = b = v[4]
// ---------------------------------------------------------------------------------------------------------------
ADV_HD void mirror_pairs(const float2* v, int l, float2* a, float2* b) {
    const bool z = (l == 0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        a[i] = v[i];
        b[i] = sel(z, v[(8 - i) & 7], v[15 - i]);
    }
#pragma unroll
    for (int i = 4; i < 8; ++i) {
        a[i] = v[4 + i];
        b[i] = sel(z, v[19 - i], v[11 - i]);
    }
    a[8] = v[4];
    b[8] = v[4];
}
ADV_HD void split(const float2* v, int l, float2* xa, float2* xb) {
    float2 a[9], b[9];
    mirror_pairs(v, l, a, b);
#pragma unroll
    for (int i = 0; i < 9; ++i) split_pair(a[i], b[i], xa[i], xb[i]);
}
// inverse placement: slot values za (bin k) and zb (bin 512 - k) -> the inverse transform's input registers.
// dc / ny: what lane 0 puts at bins 0 and 256 (C2R semantics: real parts only).
ADV_HD void place_pairs(float2* v, int l, const float2* za, const float2* zb, float2 dc, float2 ny) {
    const bool z = (l == 0);
    v[0] = sel(z, dc, za[0]);
    v[1] = za[1];
    v[2] = za[2];
    v[3] = za[3];
    v[4] = sel(z, ny, zb[7]);
    v[5] = sel(z, zb[3], zb[6]);
    v[6] = sel(z, zb[2], zb[5]);
    v[7] = sel(z, zb[1], zb[4]);
    v[8] = za[4];
    v[9] = za[5];
    v[10] = za[6];
    v[11] = za[7];
    v[12] = sel(z, zb[7], zb[3]);
    v[13] = sel(z, zb[6], zb[2]);
    v[14] = sel(z, zb[5], zb[1]);
    v[15] = sel(z, zb[4], zb[0]);
}
// two one-sided Hermitian spectra (9 slots each) -> Z = YA + i YB in the inverse transform's input layout
ADV_HD void merge(float2* v, int l, const float2* ya, const float2* yb) {
    float2 za[9], zb[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) merge_pair(ya[i], yb[i], za[i], zb[i]);
    place_pairs(v, l, za, zb, make_float2(ya[0].x, yb[0].x), make_float2(ya[8].x, yb[8].x));
}

// ---------------------------------------------------------------------------------------------------------------
// n_fft 1024 real frames through the 512-point transform: z[m] = x[2m] + i x[2m+1], Z = FFT512(z),
//   E = (Z[k] + conj Z[512-k]) / 2, O = (Z[k] - conj Z[512-k]) / (2i),  X[k] = E + W O,  X[512-k] = conj(E - W O),
//   W = exp(-2 pi i k / 1024).  Slot i of lane l yields bins k = bin_of(l, i) and 512 - k (slot 8, lane 0: bin 256
//   only; slot 0 of lane 0: bins 0 and 512).  `tw` = the shared table (TW1024_OFF section is read).
// ---------------------------------------------------------------------------------------------------------------
ADV_HD void r1024_post(const float2* v, int l, const float2* tw, float2* xk, float2* xm) {
    float2 a[9], b[9];
    mirror_pairs(v, l, a, b);
    const float2* w = tw + TW1024_OFF;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        float2 e, o;
        split_pair(a[i], b[i], e, o);
        const int k = i < 4 ? l + 64 * i : (i < 8 ? q1_of(l) + 64 * (i - 4) : 256);
        const float2 t = cmul(o, w[k]);
        xk[i] = make_float2(e.x + t.x, e.y + t.y);
        xm[i] = make_float2(e.x - t.x, t.y - e.y);
    }
}
// inverse: xk = X[k], xm = X[512-k] per slot (slot 8: both X[256]) -> inverse transform input; the result of
// fft_inverse is then 1024 * (x[2m] + i x[2m+1]) in register 2 n2 + c <-> m = l + 32 c + 64 n2.
ADV_HD void r1024_pre(float2* v, int l, const float2* tw, const float2* xk, const float2* xm) {
    float2 za[9], zb[9];
    const float2* w = tw + TW1024_OFF;
    float2 e0 = make_float2(0.f, 0.f), o0 = e0, e8 = e0, o8 = e0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const float2 e = make_float2(xk[i].x + xm[i].x, xk[i].y - xm[i].y);   // X[k] + conj X[512-k]
        const float2 d = make_float2(xk[i].x - xm[i].x, xk[i].y + xm[i].y);   // X[k] - conj X[512-k]
        const int k = i < 4 ? l + 64 * i : (i < 8 ? q1_of(l) + 64 * (i - 4) : 256);
        const float2 o = cmulc(d, w[k]);
        merge_pair(e, o, za[i], zb[i]);
        if (i == 0) { e0 = e; o0 = o; }
        if (i == 8) { e8 = e; o8 = o; }
    }
    place_pairs(v, l, za, zb, make_float2(e0.x, o0.x), make_float2(e8.x, o8.x));
}

}  // namespace f3
}  // namespace adv
