// In-register radix FFT building blocks and the lane-cooperative NF-point complex FFT
// ("unit FFT") used by the STFT / iSTFT / fused explain kernels.
//
// Design (B200, sm_100a):
//   * One *unit* = NF/32 lanes of a warp (16 lanes for NF=512, a full warp for NF=1024).
//     Every lane holds 32 complex values in registers.
//   * NF = 32 x R2 four-step FFT:  radix-32 in registers over n1 (stride R2), twiddle
//     W_NF^(n2*k1), ONE shared-memory transpose inside the unit (no block barrier, only
//     __syncwarp), radix-R2 in registers over n2.  Output index k = k1 + 32*k2.
//   * The inverse runs the mirrored data flow, so it consumes exactly the register layout
//     the forward produces and vice versa: forward -> (mask) -> inverse needs no re-layout.
//   * Row ownership after the transpose pairs row k1 with row 32-k1 in the same lane
//     (NF=512) so that bins k and NF-k are in the same thread: the two-real-frames-in-one-
//     complex-FFT split/merge is register-local.  For NF=1024 each lane owns one row and
//     the mirror bin lives in lane (32-l)&31: one shuffle per value.
//
// Everything here is __host__ __device__ so tests/host_emul.cu can run the same code
// lane-by-lane on the CPU (no GPU in the build container).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define ADV_HD __host__ __device__ __forceinline__

namespace adv {

template <int I> struct IC { static constexpr int value = I; };
template <int B, int E, class Fn>
ADV_HD void static_for(Fn&& f) {
    if constexpr (B < E) {
        f(IC<B>{});
        static_for<B + 1, E>(f);
    }
}

// cos/sin(2*pi*k/32), k = 0..15 (butterfly twiddles only need the upper half-plane)
static constexpr float kCos32[16] = {
    1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
    0.70710678118654757f, 0.55557023301960229f, 0.38268343236508984f, 0.19509032201612833f,
    0.0f, -0.19509032201612819f, -0.38268343236508973f, -0.55557023301960196f,
    -0.70710678118654746f, -0.83146961230254535f, -0.92387953251128674f, -0.98078528040323043f};
static constexpr float kSin32[16] = {
    0.0f, 0.19509032201612825f, 0.38268343236508978f, 0.55557023301960218f,
    0.70710678118654746f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f,
    1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254546f,
    0.70710678118654757f, 0.55557023301960218f, 0.38268343236508989f, 0.19509032201612861f};

// Complex add / subtract / scaled add as ONE packed instruction on sm_100a (FADD2 / FFMA2 work on an aligned
// 64-bit register pair).  Measured on B200: a packed instruction occupies the FMA pipe for two cycles, i.e. the
// same lane throughput as two scalar ones, but takes a single issue slot - and these kernels are issue-bound
// with more than half of their instructions outside the FMA pipe.  Bit-identical to the scalar forms.
#if defined(__CUDA_ARCH__) && !defined(ADV_NO_F32X2)
#define ADV_U64(v) (*reinterpret_cast<const unsigned long long*>(&(v)))
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
    float2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long*>(&r)) : "l"(ADV_U64(a)), "l"(ADV_U64(b)));
    return r;
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
    float2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long*>(&r)) : "l"(ADV_U64(a)), "l"(ADV_U64(b)));
    return r;
}
// a * (s, s) + c
__device__ __forceinline__ float2 cfma(float2 a, float2 s, float2 c) {
    float2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(*reinterpret_cast<unsigned long long*>(&r))
        : "l"(ADV_U64(a)), "l"(ADV_U64(s)), "l"(ADV_U64(c)));
    return r;
}
#else
ADV_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
ADV_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
ADV_HD float2 cfma(float2 a, float2 s, float2 c) { return make_float2(fmaf(a.x, s.x, c.x), fmaf(a.y, s.y, c.y)); }
#endif
// PK = false: plain scalar forms.  The register-starved narrow-unit inverse (istft_p_kernel holds the next tile's
// rows in flight) measured SLOWER with the packed forms (41.8 vs 39.3 us: pair alignment costs moves), every other
// kernel faster (explain 98.5 -> 94.5 us).
template <bool PK> ADV_HD float2 cadd_t(float2 a, float2 b) {
    if constexpr (PK) return cadd(a, b); else return make_float2(a.x + b.x, a.y + b.y);
}
template <bool PK> ADV_HD float2 csub_t(float2 a, float2 b) {
    if constexpr (PK) return csub(a, b); else return make_float2(a.x - b.x, a.y - b.y);
}
template <bool PK> ADV_HD float2 cfma_t(float2 a, float2 s, float2 c) {
    if constexpr (PK) return cfma(a, s, c); else return make_float2(fmaf(a.x, s.x, c.x), fmaf(a.y, s.y, c.y));
}
ADV_HD float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
ADV_HD float2 cmulc(float2 a, float2 b) {
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
ADV_HD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// a * exp(DIR * 2*pi*i * K / N), K in [0, N/2), N in {2,4,8,16,32}; DIR = -1 forward, +1 inverse
template <int N, int K, int DIR>
ADV_HD float2 twmul(float2 a) {
    static_assert(K >= 0 && 2 * K < N || N == 1, "butterfly twiddle index");
    if constexpr (K == 0) {
        return a;
    } else if constexpr (4 * K == N) {
        return DIR < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
    } else if constexpr (8 * K == N) {
        constexpr float h = 0.70710678118654752f;
        return DIR < 0 ? make_float2((a.x + a.y) * h, (a.y - a.x) * h)
                       : make_float2((a.x - a.y) * h, (a.x + a.y) * h);
    } else if constexpr (8 * K == 3 * N) {
        constexpr float h = 0.70710678118654752f;
        return DIR < 0 ? make_float2((a.y - a.x) * h, -(a.x + a.y) * h)
                       : make_float2(-(a.x + a.y) * h, (a.x - a.y) * h);
    } else {
        constexpr int idx = K * (32 / N);
        constexpr float c = kCos32[idx];
        constexpr float s = (DIR < 0 ? -1.0f : 1.0f) * kSin32[idx];
        return make_float2(a.x * c - a.y * s, a.x * s + a.y * c);
    }
}

// radix-2 butterfly with the twiddle folded into FMAs:  lo = e + W*o,  hi = e - W*o = 2e - lo
// (6 FMA-pipe instructions for a general twiddle instead of 8; trivial twiddles stay pure adds)
template <int N, int K, int DIR, bool PK = true>
ADV_HD void butterfly(float2 e, float2 o, float2& lo, float2& hi) {
    if constexpr (K == 0 || 4 * K == N) {
        const float2 t = twmul<N, K, DIR>(o);
        lo = cadd_t<PK>(e, t);
        hi = csub_t<PK>(e, t);
    } else if constexpr (8 * K == N || 8 * K == 3 * N) {
        constexpr float h = 0.70710678118654752f;
        // W*o = h * (p, q) with p, q sums / differences of o's parts
        float p, q;
        if constexpr (8 * K == N) {
            p = DIR < 0 ? o.x + o.y : o.x - o.y;
            q = DIR < 0 ? o.y - o.x : o.x + o.y;
        } else {
            p = DIR < 0 ? o.y - o.x : -(o.x + o.y);
            q = DIR < 0 ? -(o.x + o.y) : o.x - o.y;
        }
        const float2 pq = make_float2(p, q);
        lo = cfma_t<PK>(pq, make_float2(h, h), e);
        hi = cfma_t<PK>(pq, make_float2(-h, -h), e);
    } else {
        constexpr int idx = K * (32 / N);
        constexpr float c = kCos32[idx];
        constexpr float s = (DIR < 0 ? -1.0f : 1.0f) * kSin32[idx];
        lo = make_float2(fmaf(o.x, c, fmaf(-o.y, s, e.x)), fmaf(o.x, s, fmaf(o.y, c, e.y)));
        hi = cfma_t<PK>(e, make_float2(2.0f, 2.0f), make_float2(-lo.x, -lo.y));
    }
}

// N-point DFT of in[0], in[S], in[2S], ... -> out[0..N-1] (natural order), all in registers.
template <int N, int DIR, bool PK = true>
struct FFTReg {
    template <int S>
    static ADV_HD void run(const float2* in, float2* out) {
        float2 e[N / 2], o[N / 2];
        FFTReg<N / 2, DIR, PK>::template run<2 * S>(in, e);
        FFTReg<N / 2, DIR, PK>::template run<2 * S>(in + S, o);
        static_for<0, N / 2>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            butterfly<N, k, DIR, PK>(e[k], o[k], out[k], out[k + N / 2]);
        });
    }
};
template <int DIR, bool PK>
struct FFTReg<4, DIR, PK> {
    template <int S>
    static ADV_HD void run(const float2* in, float2* out) {
        const float2 a = cadd_t<PK>(in[0], in[2 * S]), b = csub_t<PK>(in[0], in[2 * S]);
        const float2 c = cadd_t<PK>(in[S], in[3 * S]);
        const float2 d = twmul<4, 1, DIR>(csub_t<PK>(in[S], in[3 * S]));
        out[0] = cadd_t<PK>(a, c);
        out[2] = csub_t<PK>(a, c);
        out[1] = cadd_t<PK>(b, d);
        out[3] = csub_t<PK>(b, d);
    }
};
template <int DIR, bool PK>
struct FFTReg<2, DIR, PK> {
    template <int S>
    static ADV_HD void run(const float2* in, float2* out) {
        out[0] = cadd_t<PK>(in[0], in[S]);
        out[1] = csub_t<PK>(in[0], in[S]);
    }
};

template <int N, int DIR, bool PK = true>
ADV_HD void fft_inplace(float2* v) {
    float2 t[N];
    FFTReg<N, DIR, PK>::template run<1>(v, t);
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = t[i];
}

// ---------------------------------------------------------------------------------------------
// Geometry of the lane-cooperative NF-point FFT
// ---------------------------------------------------------------------------------------------
template <int NF>
struct Geo {
    static_assert(NF == 512 || NF == 1024, "supported FFT sizes: 512, 1024");
    static constexpr int R2 = NF / 32;      // second-step radix == lanes per unit
    static constexpr int LANES = R2;        // lanes cooperating on one FFT
    static constexpr int ROWS = 32 / LANES; // rows (k1) owned per lane after the transpose
    static constexpr int PITCH = R2 + 1;    // scratch row pitch in float2 (odd: conflict-free)
    // floats per unit; +16 when two units share a warp so that their bank sets interleave
    static constexpr int SCRATCH = 32 * PITCH + (LANES == 16 ? 16 : 0);
    static constexpr int NBINS = NF / 2 + 1;
    static constexpr int BINS_PER_LANE = 17;  // 16 (+ Nyquist on lane 0)
};

// Row k1 owned by lane l in slot r (r < ROWS).  NF=512: {l, 32-l} (lane 0: {0, 16}); NF=1024: {l}.
template <int NF>
ADV_HD int row_of(int l, int r) {
    if constexpr (NF == 512) return r == 0 ? l : (l == 0 ? 16 : 32 - l);
    else return l;
}

// One-sided bin index handled by lane l in slot i (0..16); -1 when the slot is empty.
template <int NF>
ADV_HD int bin_of(int l, int i) {
    if constexpr (NF == 512) {
        if (i < 8) return l + 32 * i;
        if (i < 16) return (l == 0 ? 16 : 32 - l) + 32 * (i - 8);
        return l == 0 ? 256 : -1;
    } else {
        if (i < 16) return l + 32 * i;
        return l == 0 ? 512 : -1;
    }
}

// ---- transposes through shared memory, real and imaginary planes one after the other ------------
// (halves the scratch footprint: 32 x PITCH floats per unit).  Every step is followed by a unit-wide
// sync (the caller's __syncwarp); the host emulation runs each step over all lanes in turn.
template <int NF>
ADV_HD void scr_store_cols(const float2* v, int l, float* scr, bool imag) {
    using G = Geo<NF>;
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) scr[k1 * G::PITCH + l] = imag ? v[k1].y : v[k1].x;
}
template <int NF>
ADV_HD void scr_load_rows(float2* v, int l, const float* scr, bool imag) {
    using G = Geo<NF>;
#pragma unroll
    for (int r = 0; r < G::ROWS; ++r) {
        const int row = row_of<NF>(l, r);
#pragma unroll
        for (int i = 0; i < G::R2; ++i) {
            const float x = scr[row * G::PITCH + i];
            if (imag) v[r * G::R2 + i].y = x; else v[r * G::R2 + i].x = x;
        }
    }
}
template <int NF>
ADV_HD void scr_store_rows(const float2* v, int l, float* scr, bool imag) {
    using G = Geo<NF>;
#pragma unroll
    for (int r = 0; r < G::ROWS; ++r) {
        const int row = row_of<NF>(l, r);
#pragma unroll
        for (int i = 0; i < G::R2; ++i) scr[row * G::PITCH + i] = imag ? v[r * G::R2 + i].y : v[r * G::R2 + i].x;
    }
}
template <int NF>
ADV_HD void scr_load_cols(float2* v, int l, const float* scr, bool imag) {
    using G = Geo<NF>;
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) {
        const float x = scr[k1 * G::PITCH + l];
        if (imag) v[k1].y = x; else v[k1].x = x;
    }
}

// ---- same transposes with 128-bit row accesses (row pitch R2 + 4 floats: rows are 16-byte aligned and
// the eight lanes of a quarter-warp phase hit eight different 4-bank groups; column accesses stay scalar
// and conflict-free).  64 + 16 shared-memory instructions per plane pair instead of 128. -----------------
template <int NF>
struct GeoV {
    static constexpr int R2 = NF / 32;
    static constexpr int PITCH = R2 + 4;
    // +16 floats when two units share a warp: their column stores land in different bank halves
    static constexpr int SCRATCH = 32 * PITCH + (R2 == 16 ? 16 : 0);
};
template <int NF>
ADV_HD void scrv_store_cols(const float2* v, int l, float* scr, bool imag) {
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) scr[k1 * GeoV<NF>::PITCH + l] = imag ? v[k1].y : v[k1].x;
}
template <int NF>
ADV_HD void scrv_load_cols(float2* v, int l, const float* scr, bool imag) {
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) {
        const float x = scr[k1 * GeoV<NF>::PITCH + l];
        if (imag) v[k1].y = x; else v[k1].x = x;
    }
}
template <int NF>
ADV_HD void scrv_load_rows(float2* v, int l, const float* scr, bool imag) {
    using G = Geo<NF>;
#pragma unroll
    for (int r = 0; r < G::ROWS; ++r) {
        const float4* p = reinterpret_cast<const float4*>(scr + row_of<NF>(l, r) * GeoV<NF>::PITCH);
#pragma unroll
        for (int i = 0; i < G::R2 / 4; ++i) {
            const float4 q = p[i];
            float2* d = v + r * G::R2 + 4 * i;
            if (imag) { d[0].y = q.x; d[1].y = q.y; d[2].y = q.z; d[3].y = q.w; }
            else      { d[0].x = q.x; d[1].x = q.y; d[2].x = q.z; d[3].x = q.w; }
        }
    }
}
template <int NF>
ADV_HD void scrv_store_rows(const float2* v, int l, float* scr, bool imag) {
    using G = Geo<NF>;
#pragma unroll
    for (int r = 0; r < G::ROWS; ++r) {
        float4* p = reinterpret_cast<float4*>(scr + row_of<NF>(l, r) * GeoV<NF>::PITCH);
#pragma unroll
        for (int i = 0; i < G::R2 / 4; ++i) {
            const float2* s = v + r * G::R2 + 4 * i;
            p[i] = imag ? make_float4(s[0].y, s[1].y, s[2].y, s[3].y) : make_float4(s[0].x, s[1].x, s[2].x, s[3].x);
        }
    }
}

// ---- forward: time samples -> spectrum ------------------------------------------------------
// in : v[n1] = z[n1*R2 + l]          (n1 = 0..31)
// out: v[r*R2 + k2] = Z[row_of(l,r) + 32*k2]
// tw : per-lane twiddles tw(k1) = exp(-2*pi*i * l*k1 / NF), read through a functor so the caller
//      decides where they live (registers, shared memory)
template <int NF, class TwFn>
ADV_HD void fwd_cols(float2* v, TwFn tw) {  // radix-32 over n1 + twiddle
    fft_inplace<32, -1>(v);
#pragma unroll
    for (int k1 = 1; k1 < 32; ++k1) v[k1] = cmul(v[k1], tw(k1));
}
template <int NF>
ADV_HD void fwd_rows(float2* v) {  // radix-R2 over n2 for each owned row
    using G = Geo<NF>;
#pragma unroll
    for (int r = 0; r < G::ROWS; ++r) fft_inplace<G::R2, -1>(v + r * G::R2);
}
// ---- inverse: spectrum -> time samples (unnormalised) ---------------------------------------
// in : v[r*R2 + k2] = Z[row_of(l,r) + 32*k2]
// out: v[n1] = NF * z[n1*R2 + l]
template <int NF>
ADV_HD void inv_rows(float2* v) {
    using G = Geo<NF>;
#pragma unroll
    for (int r = 0; r < G::ROWS; ++r) fft_inplace<G::R2, +1, false>(v + r * G::R2);
}
template <int NF, class TwFn>
ADV_HD void inv_cols(float2* v, TwFn tw) {
#pragma unroll
    for (int k1 = 1; k1 < 32; ++k1) v[k1] = cmulc(v[k1], tw(k1));
    fft_inplace<32, +1, false>(v);
}

#ifdef __CUDACC__
// Device-side composition; `scr` is the unit's private scratch.  Both functions begin with a
// __syncwarp so that a previous transpose through the same scratch has been fully read.
template <int NF, class TwFn>
__device__ __forceinline__ void unit_fft_forward(float2* v, int l, TwFn tw, float* scr) {
    fwd_cols<NF>(v, tw);
    __syncwarp();
    scr_store_cols<NF>(v, l, scr, false);
    __syncwarp();
    scr_load_rows<NF>(v, l, scr, false);
    __syncwarp();
    scr_store_cols<NF>(v, l, scr, true);
    __syncwarp();
    scr_load_rows<NF>(v, l, scr, true);
    fwd_rows<NF>(v);
}
template <int NF, class TwFn>
__device__ __forceinline__ void unit_fft_inverse(float2* v, int l, TwFn tw, float* scr) {
    inv_rows<NF>(v);
    __syncwarp();
    scr_store_rows<NF>(v, l, scr, false);
    __syncwarp();
    scr_load_cols<NF>(v, l, scr, false);
    __syncwarp();
    scr_store_rows<NF>(v, l, scr, true);
    __syncwarp();
    scr_load_cols<NF>(v, l, scr, true);
    inv_cols<NF>(v, tw);
}
// vector-transpose variants (scratch of GeoV<NF>::SCRATCH floats per unit)
template <int NF, class TwFn>
__device__ __forceinline__ void unit_fft_forward_v(float2* v, int l, TwFn tw, float* scr) {
    fwd_cols<NF>(v, tw);
    __syncwarp();
    scrv_store_cols<NF>(v, l, scr, false);
    __syncwarp();
    scrv_load_rows<NF>(v, l, scr, false);
    __syncwarp();
    scrv_store_cols<NF>(v, l, scr, true);
    __syncwarp();
    scrv_load_rows<NF>(v, l, scr, true);
    fwd_rows<NF>(v);
}
template <int NF, class TwFn>
__device__ __forceinline__ void unit_fft_inverse_v(float2* v, int l, TwFn tw, float* scr) {
    inv_rows<NF>(v);
    __syncwarp();
    scrv_store_rows<NF>(v, l, scr, false);
    __syncwarp();
    scrv_load_cols<NF>(v, l, scr, false);
    __syncwarp();
    scrv_store_rows<NF>(v, l, scr, true);
    __syncwarp();
    scrv_load_cols<NF>(v, l, scr, true);
    inv_cols<NF>(v, tw);
}
#endif

// ---------------------------------------------------------------------------------------------
// Two real frames in one complex FFT: split (after forward) and merge (before inverse).
//   Z = FFT(xa + i*xb)  =>  XA[k] = (Z[k] + conj Z[NF-k]) / 2,  XB[k] = (Z[k] - conj Z[NF-k]) / (2i)
//   Z[k] = YA[k] + i*YB[k], Z[NF-k] = conj(YA[k]) + i*conj(YB[k])   (YA, YB Hermitian halves)
// Lane l handles the one-sided bins bin_of(l, i), i = 0..16.
// ---------------------------------------------------------------------------------------------
ADV_HD void split_pair(float2 a, float2 b, float2& xa, float2& xb) {
    // a = Z[k], b = Z[NF-k]
    xa = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
    xb = make_float2(0.5f * (a.y + b.y), 0.5f * (b.x - a.x));
}
ADV_HD void merge_pair(float2 ya, float2 yb, float2& a, float2& b) {
    // a = Z[k] = ya + i*yb ; b = Z[NF-k] = conj(ya) + i*conj(yb)
    a = make_float2(ya.x - yb.y, ya.y + yb.x);
    b = make_float2(ya.x + yb.y, yb.x - ya.y);
}
ADV_HD float2 sel(bool c, float2 a, float2 b) { return c ? a : b; }

// NF = 512, register-local.  v holds the forward output layout; xa/xb get 17 slots.
ADV_HD void split512(const float2* v, int l, float2* xa, float2* xb) {
    const bool z = (l == 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {  // bins l + 32 i (row r0); mirror in row r1 (lane 0: in row 0)
        const float2 a = v[i];
        const float2 b = sel(z, v[(16 - i) & 15], v[31 - i]);
        split_pair(a, b, xa[i], xb[i]);
    }
#pragma unroll
    for (int i = 8; i < 16; ++i) {  // bins r1 + 32 (i-8); mirror in row r0 (lane 0: in row 16)
        const float2 a = v[i + 8];
        const float2 b = sel(z, v[39 - i], v[23 - i]);
        split_pair(a, b, xa[i], xb[i]);
    }
    // Nyquist (lane 0 only, bin 256 = row 0, k2 = 8): self-mirrored
    split_pair(v[8], v[8], xa[16], xb[16]);
}
// Inverse of split512: builds the inverse-FFT input layout from two one-sided spectra.
// Imaginary parts of DC / Nyquist are dropped (C2R semantics).
ADV_HD void merge512(float2* v, int l, const float2* ya, const float2* yb) {
    const bool z = (l == 0);
    float2 za[17], zb[17];
#pragma unroll
    for (int i = 0; i < 17; ++i) merge_pair(ya[i], yb[i], za[i], zb[i]);
    const float2 dc = make_float2(ya[0].x, yb[0].x);
    const float2 ny = make_float2(ya[16].x, yb[16].x);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = za[i];
#pragma unroll
    for (int i = 8; i < 16; ++i) v[i + 8] = za[i];
    if (z) v[0] = dc;
    // mirrors: lane>0: zb[i<8] -> 31-i, zb[i>=8] -> 23-i ; lane 0: zb[i<8] -> 16-i (i>=1),
    // zb[i>=8] -> 39-i, Nyquist -> 8
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        v[24 + q] = sel(z, zb[15 - q], zb[7 - q]);
        v[8 + q] = sel(z, q == 0 ? ny : zb[8 - q], zb[15 - q]);
    }
}

// NF = 1024: the mirror of bin l+32 i lives in lane (32-l)&31 at register 31-i (lane 0: own 32-i).
// Split is 3 steps so the host emulation can perform the exchange: pre -> exchange16 -> post.
ADV_HD void split1024_pre(const float2* v, float2* send) {
#pragma unroll
    for (int i = 0; i < 16; ++i) send[i] = v[31 - i];
}
ADV_HD void split1024_post(const float2* v, int l, const float2* recv, float2* xa, float2* xb) {
    const bool z = (l == 0);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float2 b = sel(z, v[(32 - i) & 31], recv[i]);
        split_pair(v[i], b, xa[i], xb[i]);
    }
    split_pair(v[16], v[16], xa[16], xb[16]);
}
ADV_HD void merge1024_pre(float2* v, int l, const float2* ya, const float2* yb, float2* send) {
    const bool z = (l == 0);
    float2 zb[17];
#pragma unroll
    for (int i = 0; i < 16; ++i) merge_pair(ya[i], yb[i], v[i], zb[i]);
    if (z) v[0] = make_float2(ya[0].x, yb[0].x);
#pragma unroll
    for (int i = 0; i < 16; ++i) send[i] = zb[i];
    // lane 0 keeps its own mirrors: zb[i] -> register 32-i (i = 1..15), Nyquist -> 16
    if (z) {
#pragma unroll
        for (int i = 1; i < 16; ++i) v[32 - i] = zb[i];
        v[16] = make_float2(ya[16].x, yb[16].x);
    }
}
ADV_HD void merge1024_post(float2* v, int l, const float2* recv) {
    if (l != 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[31 - i] = recv[i];
    }
}

// =============================================================================================
// "Wide" 512-point unit: a full warp (32 lanes) per FFT, 16 complex values per lane.
// Half the registers per thread of the 16-lane unit => twice the resident warps for the fused explain
// kernel.  512 = 16 (in-register radix over n1, n = 32*n1 + lane) x 32 (rows): after the transpose two
// lanes share a row (r = lane>>1; lane parity h picks the even / odd n2), each runs a radix-16, and one
// butterfly across the lane pair (shuffle xor 1) finishes the 32-point row transform:
//     lane (r, h) ends with Z[r + 16*(m + 16h)], m = 0..15.
// Mirror bins (k, 512-k) live in lanes (r, 0) and (16-r, 1): the two-real-frames split / merge trades
// 8 values with that partner and every lane ends up owning 8 one-sided bins (lane 1 also the Nyquist).
// All exchange steps are split pre / post so tests/host_emul.cu can run them lane by lane.
// =============================================================================================
namespace w512 {
constexpr int LANES = 32, E = 16, PITCH = 34, SCRATCH = 16 * PITCH;  // floats of scratch per unit
constexpr int SLOTS = 9;                                             // one-sided bins per lane (8, +Nyquist on lane 1)

ADV_HD int partner_row(int l) { return 2 * ((16 - (l >> 1)) & 15) + (1 - (l & 1)); }
// one-sided bin of lane l, slot i (0..8); -1 when empty
ADV_HD int bin_of(int l, int i) {
    const int r = l >> 1, h = l & 1;
    if (i < 8) return h == 0 ? r + 16 * i : ((16 - r) & 15) + 16 * (8 + i);
    return l == 1 ? 256 : -1;
}

template <class TwFn>
ADV_HD void fwd_cols(float2* v, TwFn tw) {
    fft_inplace<16, -1>(v);
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) v[k1] = cmul(v[k1], tw(k1));
}
template <class TwFn>
ADV_HD void inv_cols(float2* v, TwFn tw) {
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) v[k1] = cmulc(v[k1], tw(k1));
    fft_inplace<16, +1>(v);
}
ADV_HD void scr_store_cols(const float2* v, int l, float* scr, bool imag) {
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) scr[k1 * PITCH + l] = imag ? v[k1].y : v[k1].x;
}
ADV_HD void scr_load_cols(float2* v, int l, const float* scr, bool imag) {
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
        const float x = scr[k1 * PITCH + l];
        if (imag) v[k1].y = x; else v[k1].x = x;
    }
}
ADV_HD void scr_load_rows(float2* v, int l, const float* scr, bool imag) {
    const float* p = scr + (l >> 1) * PITCH + (l & 1);
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const float x = p[2 * m];
        if (imag) v[m].y = x; else v[m].x = x;
    }
}
ADV_HD void scr_store_rows(const float2* v, int l, float* scr, bool imag) {
    float* p = scr + (l >> 1) * PITCH + (l & 1);
#pragma unroll
    for (int m = 0; m < 16; ++m) p[2 * m] = imag ? v[m].y : v[m].x;
}
// forward rows: radix-16 on the lane's half row, then (with `other` = partner lane's values, xor 1)
// the radix-2 across the pair
ADV_HD void fwd_rows_local(float2* v) { fft_inplace<16, -1>(v); }
ADV_HD void fwd_rows_combine(float2* v, int l, const float2* other) {
    const bool h = l & 1;
    static_for<0, 16>([&](auto mc) {
        constexpr int m = decltype(mc)::value;
        const float2 e = h ? other[m] : v[m];
        const float2 o = h ? v[m] : other[m];
        float2 lo, hi;
        butterfly<32, m, -1>(e, o, lo, hi);
        v[m] = h ? hi : lo;
    });
}
// inverse rows: undo the pair butterfly (needs the partner's values), then radix-16
ADV_HD void inv_rows_combine(float2* v, int l, const float2* other) {
    const bool h = l & 1;
    static_for<0, 16>([&](auto mc) {
        constexpr int m = decltype(mc)::value;
        // h = 0: X[m] + X[m+16];  h = 1: (X[m] - X[m+16]) * W_32^{-m}
        const float2 lo = h ? other[m] : v[m];
        const float2 hi = h ? v[m] : other[m];
        v[m] = h ? twmul<32, m, +1>(csub(lo, hi)) : cadd(lo, hi);
    });
}
ADV_HD void inv_rows_local(float2* v) { fft_inplace<16, +1>(v); }

// Lean pair butterfly (same results as fwd_rows_combine / inv_rows_combine up to round-off, a third of the
// instructions: no per-value selects).  Forward: the odd lane multiplies its half-row transform by W_32^m
// BEFORE the exchange, then both lanes do  v = other + sg * v  (sg = +1 even lane, -1 odd lane).
// Inverse: the same butterfly first, then the odd lane multiplies by W_32^{-m}.
ADV_HD void fwd_rows_tw(float2* v, int l) {
    if (l & 1) {
        static_for<1, 16>([&](auto mc) {
            constexpr int m = decltype(mc)::value;
            v[m] = twmul<32, m, -1>(v[m]);
        });
    }
}
ADV_HD void rows_bfly(float2* v, int l, const float2* other) {
    const float sg = (l & 1) ? -1.0f : 1.0f;
    const float2 sg2 = make_float2(sg, sg);
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = cfma(v[m], sg2, other[m]);
}
ADV_HD void inv_rows_tw(float2* v, int l) {
    if (l & 1) {
        static_for<1, 16>([&](auto mc) {
            constexpr int m = decltype(mc)::value;
            v[m] = twmul<32, m, +1>(v[m]);
        });
    }
}

// ---- split: Z (two packed real frames) -> 9 one-sided bins of each frame per lane ---------------------
ADV_HD void split_pre(const float2* v, float2* send) {
#pragma unroll
    for (int j = 0; j < 8; ++j) send[j] = v[8 + j];
}
ADV_HD void split_post(const float2* v, int l, const float2* recv, float2* xa, float2* xb) {
    const bool h = l & 1, z = (l >> 1) == 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float2 a, b;
        if (!h) {  // lane (r,0): bins r + 16 i; mirror came from the partner
            a = v[i];
            b = z ? (i == 0 ? v[0] : recv[8 - i]) : recv[7 - i];
        } else {   // lane (r,1): bins of the partner's row, k2 = 8 + i; own registers hold the mirrors
            a = recv[i];
            b = z ? v[8 - i] : v[7 - i];
        }
        split_pair(a, b, xa[i], xb[i]);
    }
    split_pair(v[0], v[0], xa[8], xb[8]);  // Nyquist: meaningful on lane 1 only
}
// ---- merge: two one-sided spectra -> inverse-FFT input layout (C2R semantics on DC / Nyquist) -------
ADV_HD void merge_pre(float2* v, int l, const float2* ya, const float2* yb, float2* send) {
    const bool h = l & 1, z = (l >> 1) == 0;
    float2 za[9], zb[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) merge_pair(ya[i], yb[i], za[i], zb[i]);
    if (!h) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = za[i];
        if (z) v[0] = make_float2(ya[0].x, yb[0].x);  // DC
#pragma unroll
        for (int j = 0; j < 8; ++j) send[j] = z ? zb[(8 - j) & 7] : zb[7 - j];  // (z: j = 0 is unused)
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (z) v[8 - i] = zb[i]; else v[7 - i] = zb[i];
        }
        if (z) v[0] = make_float2(ya[8].x, yb[8].x);  // Nyquist
#pragma unroll
        for (int j = 0; j < 8; ++j) send[j] = za[j];
    }
}
ADV_HD void merge_post(float2* v, int l, const float2* recv) {
    const bool keep8 = (l == 1);  // lane (0,1) computed its own v[8] (mirror of its k2 = 8 bin)
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (!(keep8 && j == 0)) v[8 + j] = recv[j];
}
}  // namespace w512

}  // namespace adv
