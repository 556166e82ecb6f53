// Tensor-core (tcgen05 / TMEM) kernels of the path, sm_100a:
//   conv1d_tc_kernel  - bf16 channels-last conv1d / phase-decomposed transposed conv as an implicit GEMM
//                       (HiFi-GAN generator that hifigan.py:180 runs through SpeechBrain's decode_batch)
//   mel_tc_kernel     - |X|^p x mel filterbank as a 3xTF32 split GEMM with a log-compress epilogue
//                       (audioprocessor.py:38-44 MelSpectrogram; hifigan.py:163-178 mel_spectogram)
// plus the small CUDA-core kernels around them (layout conversion, MRF average, 1-channel post conv).
//
// GEMM view: D[M = 128 rows per CTA][N] += A[M][K] * B[N][K]^T, fp32 accumulation in TMEM.
//   conv: row m = (clip b, position l); K index = tap * C_in + c_in; A is gathered on the fly from the
//         channels-last activation (zero / reflect padding, optional LeakyReLU on load), B = weights.
//   mel : row m = (clip b, frame t); K = frequency bin; A = |X|^p computed on the fly from the complex
//         spectrum, split into tf32 hi + lo parts; B = filterbank^T split the same way (hi*hi + lo*hi + hi*lo).
// Operands are written to shared memory in the canonical K-major SWIZZLE_128B layout (umma.cuh), two
// stages deep: while the tensor core works on stage s the CTA's 128 threads gather stage s^1.
#include <cuda_bf16.h>
#include <mutex>
#include <unordered_map>
#include "adv_internal.cuh"
#include "umma.cuh"

namespace adv {

using namespace umma;

constexpr int kGemmThreads = 128;
constexpr int kTileM = 128;
constexpr int kRowBytes = 128;                 // one swizzled operand row
constexpr int kATileBytes = kTileM * kRowBytes;  // 16 KB

struct ConvArgs {
    const __nv_bfloat16* in;     // [B][L][Cin]
    const __nv_bfloat16* w;      // [N][Kpad], k = tap*Cin + ci
    const float* bias;           // [N] or null
    const __nv_bfloat16* resid;  // [B][L][N] or null
    __nv_bfloat16* out;          // [B][L][N] or null
    __nv_bfloat16* out_act;      // [B][L][N] or null: LeakyReLU(act_slope) of the result
    float act_slope;
    int B, L, Cin, taps, dil, center, N, Kreal, Kpad, pad_reflect;
    float pre_slope;             // LeakyReLU slope applied to the input on load (1 = identity)
    float out_scale;
};

__device__ __forceinline__ uint32_t lrelu_bf16x2(uint32_t x, __nv_bfloat162 slope) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&x);
    __nv_bfloat162 r = __hmax2(v, __hmul2(v, slope));  // slope in (0,1): max(x, slope*x)
    return *reinterpret_cast<uint32_t*>(&r);
}

template <int BN>
__global__ void __launch_bounds__(kGemmThreads)
conv1d_tc_kernel(ConvArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kBTileBytes = BN * kRowBytes;
    constexpr int kStageBytes = kATileBytes + kBTileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kStageBytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long M = (long)a.B * a.L;
    const long m0 = (long)blockIdx.x * kTileM;
    const int n0 = blockIdx.y * BN;

    if (tid == 0) {
        bar_init(&bars[0], 1);
        bar_init(&bars[1], 1);
        bar_init_fence();
    }
    if (warp == 0) tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_d = *tmem_slot;

    // rows this thread gathers: (tid >> 3) + 16 i, chunk kc = tid & 7 (8 threads cover one 128-byte row)
    const int kc = tid & 7;
    int row_b[8], row_l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long m = m0 + (tid >> 3) + 16 * i;
        if (m < M) {
            row_b[i] = (int)(m / a.L);
            row_l[i] = (int)(m - (long)row_b[i] * a.L);
        } else {
            row_b[i] = -1;
            row_l[i] = 0;
        }
    }
    const bool act_in = a.pre_slope != 1.0f;
    const __nv_bfloat162 slope2 = __float2bfloat162_rn(a.pre_slope);
    constexpr uint32_t idesc = make_idesc(FMT_BF16, kTileM, BN);
    const int nkb = a.Kpad / 64;

    for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb & 1;
        unsigned char* sA = smem + s * kStageBytes;
        unsigned char* sB = sA + kATileBytes;
        if (kb >= 2) bar_wait(&bars[s], ((kb >> 1) - 1) & 1);  // the MMAs that read this stage have retired
        // ---- gather A: 128 rows x 64 channels-of-K ----
        const int kg = kb * 64 + kc * 8;
        const int tap = kg / a.Cin, ci = kg - tap * a.Cin;
        const int shift = (tap - a.center) * a.dil;
        const bool k_ok = kg < a.Kreal;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int4 val = make_int4(0, 0, 0, 0);
            if (k_ok && row_b[i] >= 0) {
                int sl = row_l[i] + shift;
                if (a.pad_reflect) {
                    if (sl < 0) sl = -sl;
                    else if (sl >= a.L) sl = 2 * (a.L - 1) - sl;
                }
                if (sl >= 0 && sl < a.L) {
                    val = __ldg(reinterpret_cast<const int4*>(a.in + ((size_t)row_b[i] * a.L + sl) * a.Cin + ci));
                    if (act_in) {
                        val.x = lrelu_bf16x2(val.x, slope2);
                        val.y = lrelu_bf16x2(val.y, slope2);
                        val.z = lrelu_bf16x2(val.z, slope2);
                        val.w = lrelu_bf16x2(val.w, slope2);
                    }
                }
            }
            *reinterpret_cast<int4*>(sA + sw128_offset((tid >> 3) + 16 * i, kc)) = val;
        }
        // ---- gather B: BN rows of weights ----
#pragma unroll
        for (int i = 0; i < BN / 16; ++i) {
            const int row = (tid >> 3) + 16 * i;
            const int4 val = __ldg(reinterpret_cast<const int4*>(a.w + (size_t)(n0 + row) * a.Kpad + kb * 64 + kc * 8));
            *reinterpret_cast<int4*>(sB + sw128_offset(row, kc)) = val;
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            fence_after_sync();
            const uint64_t da = make_desc_sw128(smem_addr(sA));
            const uint64_t db = make_desc_sw128(smem_addr(sB));
#pragma unroll
            for (int k = 0; k < 4; ++k)  // 4 x (K = 16 bf16 = 32 bytes): advance the start address by 2 (x16 B)
                mma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            mma_commit(&bars[s]);
        }
    }
    {
        const int last = nkb - 1;
        bar_wait(&bars[last & 1], (last >> 1) & 1);
        fence_after_sync();
    }

    // ---- epilogue: TMEM -> registers -> (+bias, +residual, *scale) -> bf16 channels-last ----
    const long m = m0 + warp * 32 + lane;
    const uint32_t trow = tmem_d + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 8) {
        float v[8];
        tmem_ld8(trow + c0, v);
        if (m < M) {
            const int n = n0 + c0;
            if (a.bias != nullptr) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] += __ldg(a.bias + n + j);
            }
            if (a.resid != nullptr) {
                const int4 r = __ldg(reinterpret_cast<const int4*>(a.resid + (size_t)m * a.N + n));
                const __nv_bfloat162* rp = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __bfloat1622float2(rp[j]);
                    v[2 * j] += f.x;
                    v[2 * j + 1] += f.y;
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] *= a.out_scale;
            if (a.out != nullptr) {
                int4 o;
                __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                for (int j = 0; j < 4; ++j) op[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                *reinterpret_cast<int4*>(a.out + (size_t)m * a.N + n) = o;
            }
            if (a.out_act != nullptr) {
                int4 o;
                __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float x0 = v[2 * j], x1 = v[2 * j + 1];
                    op[j] = __floats2bfloat162_rn(x0 > 0.f ? x0 : x0 * a.act_slope, x1 > 0.f ? x1 : x1 * a.act_slope);
                }
                *reinterpret_cast<int4*>(a.out_act + (size_t)m * a.N + n) = o;
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, BN < 32 ? 32 : BN);
}

// ---------------------------------------------------------------------------------------------------
// mel projection: out[b][mel][t] = post( sum_f fb[f][mel] * |X[b][t][f]|^power )
// ---------------------------------------------------------------------------------------------------
struct MelArgs {
    const float2* X;     // [B*T][F] frame-major spectrum
    const float* fb_hi;  // [NM][Kpad] filterbank^T, tf32-truncated part
    const float* fb_lo;  // [NM][Kpad] remainder
    float* out;          // [B][n_mels][T]
    long rows;           // B*T
    int T, F, Kpad, n_mels;
    float power;         // 1 or 2 (anything else goes through powf)
    int log_compress;    // log(clamp(x, clip))
    float clip;
};

__device__ __noinline__ float mel_pow_slow(float r2, float e) { return powf(r2, e); }

// 512 threads: every K-block's gather + tf32 split is spread over 16 warps (8 bins per thread instead of 32), which
// is what hides the global-load latency with one CTA per SM; warps 0-3 own the TMEM lanes for the epilogue.
constexpr int kMelThreads = 512;

template <int NM>  // padded mel count (multiple of 16)
__global__ void __launch_bounds__(kMelThreads)
mel_tc_kernel(MelArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kBTile = ((NM * kRowBytes + 1023) / 1024) * 1024;
    constexpr int kStageBytes = 2 * kATileBytes + 2 * kBTile;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kStageBytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    constexpr uint32_t kCols = NM <= 32 ? 32 : (NM <= 64 ? 64 : (NM <= 128 ? 128 : 256));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long m0 = (long)blockIdx.x * kTileM;
    if (tid == 0) {
        bar_init(&bars[0], 1);
        bar_init(&bars[1], 1);
        bar_init_fence();
    }
    if (warp == 0) tmem_alloc(tmem_slot, kCols);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_d = *tmem_slot;

    const int kc = tid & 7;  // 16-byte chunk = 4 fp32 = 4 bins
    constexpr uint32_t idesc = make_idesc(FMT_TF32, kTileM, NM);
    const int nkb = a.Kpad / 32;
    constexpr int kBIters = (NM * 8 + kMelThreads - 1) / kMelThreads;
    constexpr int kRowPass = kMelThreads / 8, kAIters = kTileM / kRowPass;  // rows gathered per pass, passes

    // Two register sets: ALL global loads of K-block kb+1 (32 spectrum bins and this thread's filterbank chunks) are
    // requested before K-block kb is converted and multiplied, so their latency is covered by the conversion
    // arithmetic and the barrier instead of being paid eight times per block (the first version loaded four bins at a
    // time with 48 registers: 342 us for 64 clips, all of it exposed latency with one 128-thread CTA per SM).
    struct Regs {
        float2 xv[kAIters][4];
        float4 bh[kBIters], bl[kBIters];
    };
    auto load_block = [&](int kb, Regs& r) {
        const int f0 = kb * 32 + kc * 4;
#pragma unroll
        for (int i = 0; i < kAIters; ++i) {
            const long m = m0 + (tid >> 3) + kRowPass * i;
            const float2* src = a.X + (size_t)m * a.F + f0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                r.xv[i][j] = (m < a.rows && f0 + j < a.F) ? __ldg(src + j) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int it = 0; it < kBIters; ++it) {
            const int q = tid + it * kMelThreads;
            if (q < NM * 8) {
                const size_t g = (size_t)(q >> 3) * a.Kpad + kb * 32 + (q & 7) * 4;
                r.bh[it] = __ldg(reinterpret_cast<const float4*>(a.fb_hi + g));
                r.bl[it] = __ldg(reinterpret_cast<const float4*>(a.fb_lo + g));
            }
        }
    };
    auto do_block = [&](int kb, const Regs& r) {
        const int s = kb & 1;
        unsigned char* Ahi = smem + s * kStageBytes;
        unsigned char* Alo = Ahi + kATileBytes;
        unsigned char* Bhi = Alo + kATileBytes;
        unsigned char* Blo = Bhi + kBTile;
        if (kb >= 2) bar_wait(&bars[s], ((kb >> 1) - 1) & 1);
#pragma unroll
        for (int i = 0; i < kAIters; ++i) {
            const int row = (tid >> 3) + kRowPass * i;
            float hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 x = r.xv[i][j];
                const float r2 = fmaf(x.x, x.x, x.y * x.y);
                // power 1: r2 * rsqrt(r2) (<= 2 ulp; IEEE sqrtf's inlined slow path made the unrolled kernel 140 KB of
                // code and a third of its stall samples were instruction-cache misses); other exponents: out of line
                const float p = a.power == 2.0f ? r2
                                : (a.power == 1.0f ? r2 * rsqrtf(fmaxf(r2, 1e-37f)) : mel_pow_slow(r2, 0.5f * a.power));
                hi[j] = __uint_as_float(__float_as_uint(p) & 0xFFFFE000u);  // what the tf32 datapath keeps
                lo[j] = p - hi[j];
            }
            const uint32_t off = sw128_offset(row, kc);
            *reinterpret_cast<float4*>(Ahi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(Alo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
#pragma unroll
        for (int it = 0; it < kBIters; ++it) {
            const int q = tid + it * kMelThreads;
            if (q < NM * 8) {
                const uint32_t off = sw128_offset(q >> 3, q & 7);
                *reinterpret_cast<float4*>(Bhi + off) = r.bh[it];
                *reinterpret_cast<float4*>(Blo + off) = r.bl[it];
            }
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            fence_after_sync();
            const uint64_t dah = make_desc_sw128(smem_addr(Ahi)), dal = make_desc_sw128(smem_addr(Alo));
            const uint64_t dbh = make_desc_sw128(smem_addr(Bhi)), dbl = make_desc_sw128(smem_addr(Blo));
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // K = 8 tf32 = 32 bytes per MMA
                mma_tf32(tmem_d, dal + 2 * k, dbh + 2 * k, idesc, (kb | k) != 0);  // small terms first
                mma_tf32(tmem_d, dah + 2 * k, dbl + 2 * k, idesc, 1);
                mma_tf32(tmem_d, dah + 2 * k, dbh + 2 * k, idesc, 1);
            }
            mma_commit(&bars[s]);
        }
    };
    Regs r0, r1;
    load_block(0, r0);
    for (int kb = 0; kb < nkb; kb += 2) {
        if (kb + 1 < nkb) load_block(kb + 1, r1);
        do_block(kb, r0);
        if (kb + 1 < nkb) {
            if (kb + 2 < nkb) load_block(kb + 2, r0);
            do_block(kb + 1, r1);
        }
    }
    {
        const int last = nkb - 1;
        bar_wait(&bars[last & 1], (last >> 1) & 1);
        fence_after_sync();
    }
    const long m = m0 + warp * 32 + lane;
    const uint32_t trow = tmem_d + ((uint32_t)(warp * 32) << 16);
    const long b = m < a.rows ? m / a.T : 0;
    const int t = (int)(m - b * a.T);
#pragma unroll 1
    for (int c0 = 0; c0 < (warp < 4 ? NM : 0); c0 += 8) {
        float v[8];
        tmem_ld8(trow + c0, v);
        if (m < a.rows) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int mel = c0 + j;
                if (mel < a.n_mels) {
                    float y = v[j];
                    if (a.log_compress) y = logf(fmaxf(y, a.clip));
                    a.out[((size_t)b * a.n_mels + mel) * a.T + t] = y;  // lanes = consecutive t: coalesced
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, kCols);
}

// ---------------------------------------------------------------------------------------------------
// CUDA-core helpers of the vocoder
// ---------------------------------------------------------------------------------------------------
// mel [B][C][T] fp32 -> channels-last bf16 [B][T + 2 pad][Cpad] with replicate padding in time
__global__ void mel_to_cl_kernel(const float* __restrict__ mel, int B, int C, int T, int pad, int Cpad,
                                 __nv_bfloat16* __restrict__ out) {
    const int L = T + 2 * pad;
    const long total = (long)B * L * Cpad;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cpad);
        const long bl = i / Cpad;
        const int l = (int)(bl % L), b = (int)(bl / L);
        int t = l - pad;
        t = t < 0 ? 0 : (t >= T ? T - 1 : t);
        out[i] = __float2bfloat16(c < C ? mel[((size_t)b * C + c) * T + t] : 0.0f);
    }
}
// out = (a + b + c) / 3 (MRF average, HifiganGenerator.forward)
__global__ void avg3_kernel(const __nv_bfloat162* __restrict__ a, const __nv_bfloat162* __restrict__ b,
                            const __nv_bfloat162* __restrict__ c, long n2, float slope, __nv_bfloat162* __restrict__ out) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long)gridDim.x * blockDim.x) {
        const float2 x = __bfloat1622float2(a[i]), y = __bfloat1622float2(b[i]), z = __bfloat1622float2(c[i]);
        // the mean is rounded to bf16 first (it is the tensor the reference would hold), then activated
        const float2 m = __bfloat1622float2(__floats2bfloat162_rn((x.x + y.x + z.x) * (1.0f / 3.0f), (x.y + y.y + z.y) * (1.0f / 3.0f)));
        out[i] = __floats2bfloat162_rn(m.x > 0.f ? m.x : m.x * slope, m.y > 0.f ? m.y : m.y * slope);
    }
}
// ---- halo layout for REFLECT "same" padding on the TMA conv kernels (SpeechBrain's Conv1d default, which the
// reference's HIFIGAN uses for conv_pre, the ResBlock convs and conv_post).  An activation is stored as
// [B][L + 2 H][C] with its L real rows at offset H; the conv kernels run over all L + 2 H rows with their zero (TMA
// out-of-bounds) padding - what they compute in the halo rows is never used - and the kernel below rewrites the halo
// rows of a produced tensor IN PLACE with the reflection of its interior (mode 1) or with zeros (mode 0, the input of a
// transposed conv) before a conv reads it: interior results then see exactly F.pad(x, mode="reflect").
__global__ void halo_fix_kernel(__nv_bfloat16* __restrict__ buf, int B, int L, int C8, int H, int mode) {
    const long total = (long)B * 2 * H * C8;   // 16-byte groups in the halo rows
    int4* p = reinterpret_cast<int4*>(buf);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C8);
        const long r = i / C8;
        const int hr = (int)(r % (2 * H)), b = (int)(r / (2 * H));
        // halo row hr < H: position -(H - hr) .. -1; else position L + (hr - H)
        const int pos = hr < H ? hr - H : L + (hr - H);
        int src = pos < 0 ? -pos : 2 * (L - 1) - pos;
        const size_t row0 = (size_t)b * (L + 2 * H);
        int4 v = make_int4(0, 0, 0, 0);
        if (mode == 1 && src >= 0 && src < L) v = p[(row0 + H + src) * C8 + c];
        p[(row0 + H + pos) * C8 + c] = v;
    }
}
// out interior = LeakyReLU_slope(bf16(mean of the n_in inputs' interiors)); inputs [B][L + 2 h_in][C], output
// [B][L + 2 h_out][C] with zero (mode 0) or reflected (mode 1) halo rows: the MRF average and the change of halo width
// between generator stages in one pass
__global__ void avg_relayout_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                    const __nv_bfloat16* __restrict__ c, int B, int L, int C2, int h_in, int h_out, int mode,
                                    float slope, __nv_bfloat16* __restrict__ out) {
    const int Lo = L + 2 * h_out, Li = L + 2 * h_in;
    const long total = (long)B * Lo * C2;
    const __nv_bfloat162 *a2 = reinterpret_cast<const __nv_bfloat162*>(a), *b2 = reinterpret_cast<const __nv_bfloat162*>(b),
                         *c2 = reinterpret_cast<const __nv_bfloat162*>(c);
    __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(out);
    const float inv = c2 ? 1.0f / 3.0f : (b2 ? 0.5f : 1.0f);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % C2);
        const long r = i / C2;
        const int row = (int)(r % Lo), bb = (int)(r / Lo);
        int pos = row - h_out;
        if (pos < 0 || pos >= L) {
            if (mode == 0) { o2[i] = __floats2bfloat162_rn(0.f, 0.f); continue; }
            pos = pos < 0 ? -pos : 2 * (L - 1) - pos;
            if (pos < 0 || pos >= L) { o2[i] = __floats2bfloat162_rn(0.f, 0.f); continue; }
        }
        const size_t src = ((size_t)bb * Li + h_in + pos) * C2 + ch;
        float2 x = __bfloat1622float2(a2[src]);
        if (b2) { const float2 y = __bfloat1622float2(b2[src]); x.x += y.x; x.y += y.y; }
        if (c2) { const float2 z = __bfloat1622float2(c2[src]); x.x += z.x; x.y += z.y; }
        // the mean is rounded to bf16 first (it is the tensor the reference would hold), then activated
        const float2 m = c2 || b2 ? __bfloat1622float2(__floats2bfloat162_rn(x.x * inv, x.y * inv)) : x;
        o2[i] = __floats2bfloat162_rn(m.x > 0.f ? m.x : m.x * slope, m.y > 0.f ? m.y : m.y * slope);
    }
}
// the same with 16-byte accesses (C % 8 == 0, every production width): 8 channels per thread and step.  The 4-byte form
// above reached 42 % of the DRAM throughput on the 1.1 GB tensors of the 128-, 64- and 32-channel stages (4.1 ms of a 65 ms
// generator pass at 256 clips, profiles/r02zb_vocoder_launches_b256.csv).
__global__ void avg_relayout8_kernel(const int4* __restrict__ a, const int4* __restrict__ b, const int4* __restrict__ c, int B,
                                     int L, int C8, int h_in, int h_out, int mode, float slope, int4* __restrict__ out) {
    const int Lo = L + 2 * h_out, Li = L + 2 * h_in;
    const long total = (long)B * Lo * C8;
    const float inv = c ? 1.0f / 3.0f : (b ? 0.5f : 1.0f);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % C8);
        const long r = i / C8;
        const int row = (int)(r % Lo), bb = (int)(r / Lo);
        int pos = row - h_out;
        bool zero = false;
        if (pos < 0 || pos >= L) {
            if (mode == 0) zero = true;
            else {
                pos = pos < 0 ? -pos : 2 * (L - 1) - pos;
                zero = pos < 0 || pos >= L;
            }
        }
        int4 o = make_int4(0, 0, 0, 0);
        if (!zero) {
            const size_t src = ((size_t)bb * Li + h_in + pos) * C8 + ch;
            const int4 xa = __ldg(a + src);
            const int4 xb = b ? __ldg(b + src) : make_int4(0, 0, 0, 0);
            const int4 xc = c ? __ldg(c + src) : make_int4(0, 0, 0, 0);
            const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&xa);
            const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&xb);
            const __nv_bfloat162* pc = reinterpret_cast<const __nv_bfloat162*>(&xc);
            __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 x = __bfloat1622float2(pa[j]);
                if (b) { const float2 y = __bfloat1622float2(pb[j]); x.x += y.x; x.y += y.y; }
                if (c) { const float2 z = __bfloat1622float2(pc[j]); x.x += z.x; x.y += z.y; }
                // the mean is rounded to bf16 first (it is the tensor the reference would hold), then activated
                const float2 m = c || b ? __bfloat1622float2(__floats2bfloat162_rn(x.x * inv, x.y * inv)) : x;
                po[j] = __floats2bfloat162_rn(m.x > 0.f ? m.x : m.x * slope, m.y > 0.f ? m.y : m.y * slope);
            }
        }
        out[i] = o;
    }
}
// conv_post: LeakyReLU(slope) -> conv1d(C -> 1, taps) -> tanh; in channels-last bf16, out fp32 [B][L]
template <int C, int TAPS>
__global__ void post_conv_kernel(const __nv_bfloat16* __restrict__ in, const float* __restrict__ w /*[TAPS][C]*/,
                                 const float* __restrict__ bias, int B, int L, float slope, int pad_reflect,
                                 float* __restrict__ out) {
    __shared__ float ws[TAPS * C];
    for (int i = threadIdx.x; i < TAPS * C; i += blockDim.x) ws[i] = w[i];
    __syncthreads();
    const long total = (long)B * L;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int b = (int)(i / L), l = (int)(i - (long)b * L);
        float acc = bias[0];
#pragma unroll
        for (int tap = 0; tap < TAPS; ++tap) {
            int sl = l + tap - TAPS / 2;
            if (pad_reflect) {
                if (sl < 0) sl = -sl;
                else if (sl >= L) sl = 2 * (L - 1) - sl;
            }
            if (sl < 0 || sl >= L) continue;
            const int4* p = reinterpret_cast<const int4*>(in + ((size_t)b * L + sl) * C);
#pragma unroll
            for (int q = 0; q < C / 8; ++q) {
                const int4 r = __ldg(p + q);
                const __nv_bfloat162* rp = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float2 f = __bfloat1622float2(rp[j]);
                    f.x = f.x > 0.f ? f.x : f.x * slope;
                    f.y = f.y > 0.f ? f.y : f.y * slope;
                    acc = fmaf(f.x, ws[tap * C + q * 8 + 2 * j], acc);
                    acc = fmaf(f.y, ws[tap * C + q * 8 + 2 * j + 1], acc);
                }
            }
        }
        out[i] = tanhf(acc);
    }
}

// the same, four consecutive outputs per thread: the TAPS + 3 input rows they share are loaded and activated once
// (10 x 32 conversions for 4 outputs instead of 28 x 32), weights come from shared memory as float4.  Zero or reflect padding.
template <int C, int TAPS>
__global__ void __launch_bounds__(256)
post_conv4_kernel(const __nv_bfloat16* __restrict__ in, const float* __restrict__ w /*[TAPS][C]*/,
                  const float* __restrict__ bias, int B, int L, float slope, int pad_reflect, float* __restrict__ out) {
    __shared__ __align__(16) float ws[TAPS * C];
    for (int i = threadIdx.x; i < TAPS * C; i += blockDim.x) ws[i] = w[i];
    __syncthreads();
    constexpr int R = TAPS + 3, H = TAPS / 2;
    const int groups = (L + 3) / 4;
    const long total = (long)B * groups;
    const float b0 = bias[0];
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int b = (int)(i / groups), l0 = (int)(i - (long)b * groups) * 4;
        float acc[4] = {b0, b0, b0, b0};
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int sl = l0 + r - H;
            if (pad_reflect) {
                if (sl < 0) sl = -sl;
                else if (sl >= L) sl = 2 * (L - 1) - sl;
            }
            if (sl < 0 || sl >= L) continue;   // zero padding / outside the reflection range
            const int4* p = reinterpret_cast<const int4*>(in + ((size_t)b * L + sl) * C);
#pragma unroll
            for (int q = 0; q < C / 8; ++q) {
                const int4 raw = __ldg(p + q);
                const __nv_bfloat162* rp = reinterpret_cast<const __nv_bfloat162*>(&raw);
                float x[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __bfloat1622float2(rp[j]);
                    x[2 * j] = f.x > 0.f ? f.x : f.x * slope;
                    x[2 * j + 1] = f.y > 0.f ? f.y : f.y * slope;
                }
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    const int tap = r - o;   // row r feeds output l0 + o through tap r - o
                    if (tap < 0 || tap >= TAPS) continue;
                    const float4 w0 = *reinterpret_cast<const float4*>(ws + tap * C + q * 8);
                    const float4 w1 = *reinterpret_cast<const float4*>(ws + tap * C + q * 8 + 4);
                    float a = acc[o];
                    a = fmaf(x[0], w0.x, a); a = fmaf(x[1], w0.y, a); a = fmaf(x[2], w0.z, a); a = fmaf(x[3], w0.w, a);
                    a = fmaf(x[4], w1.x, a); a = fmaf(x[5], w1.y, a); a = fmaf(x[6], w1.z, a); a = fmaf(x[7], w1.w, a);
                    acc[o] = a;
                }
            }
        }
        float* orow = out + (size_t)b * L + l0;
        if (l0 + 3 < L && (reinterpret_cast<uintptr_t>(orow) & 15) == 0) {
            *reinterpret_cast<float4*>(orow) = make_float4(tanhf(acc[0]), tanhf(acc[1]), tanhf(acc[2]), tanhf(acc[3]));
        } else {
#pragma unroll
            for (int o = 0; o < 4; ++o)
                if (l0 + o < L) orow[o] = tanhf(acc[o]);
        }
    }
}

template <class K>
static int set_smem_attr(K kernel, size_t bytes) {
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

template <int BN>
static int launch_conv(const ConvArgs& a, cudaStream_t s) {
    const size_t smem = 2 * (kATileBytes + (size_t)BN * kRowBytes) + 64 + 1024;
    int rc = set_smem_attr(conv1d_tc_kernel<BN>, smem);
    if (rc != ADV_OK) return rc;
    const long M = (long)a.B * a.L;
    dim3 grid((unsigned)((M + kTileM - 1) / kTileM), a.N / BN);
    conv1d_tc_kernel<BN><<<grid, kGemmThreads, smem, s>>>(a);
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

}  // namespace adv

using namespace adv;

extern "C" {

int adv_conv1d_bf16(const void* in, const void* w, const float* bias, const void* resid, void* out, void* out_act,
                    int batch, int L, int Cin, int taps, int dil, int N, int Kpad, int pad_reflect, float pre_slope,
                    float act_slope, float out_scale, void* stream) {
    if (!in || !w || (!out && !out_act) || batch <= 0 || L <= 0 || Cin <= 0 || taps <= 0 || dil <= 0 || N <= 0) return ADV_ERR_INVALID;
    if (Cin % 8 != 0 || N % 16 != 0 || Kpad % 64 != 0 || Kpad < taps * Cin || (taps & 1) == 0) return ADV_ERR_SHAPE;
    ConvArgs a;
    a.in = (const __nv_bfloat16*)in;
    a.w = (const __nv_bfloat16*)w;
    a.bias = bias;
    a.resid = (const __nv_bfloat16*)resid;
    a.out = (__nv_bfloat16*)out;
    a.out_act = (__nv_bfloat16*)out_act;
    a.act_slope = act_slope;
    a.B = batch; a.L = L; a.Cin = Cin; a.taps = taps; a.dil = dil; a.center = (taps - 1) / 2;
    a.N = N; a.Kreal = taps * Cin; a.Kpad = Kpad; a.pad_reflect = pad_reflect;
    a.pre_slope = pre_slope; a.out_scale = out_scale;
    cudaStream_t s = (cudaStream_t)stream;
    if (N % 256 == 0) return launch_conv<256>(a, s);
    if (N % 128 == 0) return launch_conv<128>(a, s);
    if (N % 64 == 0) return launch_conv<64>(a, s);
    if (N % 32 == 0) return launch_conv<32>(a, s);
    return launch_conv<16>(a, s);
}

int adv_mel_project(const adv_c64* X, int64_t rows, int T, int F, const float* fb_hi, const float* fb_lo, int Kpad,
                    int n_mels, float power, int log_compress, float clip, float* out, void* stream) {
    if (!X || !fb_hi || !fb_lo || !out || rows <= 0 || T <= 0 || F <= 0 || n_mels <= 0) return ADV_ERR_INVALID;
    if (Kpad % 32 != 0 || Kpad < F || n_mels > 128) return ADV_ERR_SHAPE;
    MelArgs a;
    a.X = (const float2*)X; a.fb_hi = fb_hi; a.fb_lo = fb_lo; a.out = out; a.rows = rows; a.T = T; a.F = F;
    a.Kpad = Kpad; a.n_mels = n_mels; a.power = power; a.log_compress = log_compress; a.clip = clip;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((rows + kTileM - 1) / kTileM);
    int rc;
#define ADV_MEL(NM)                                                                                     \
    do {                                                                                                \
        const size_t smem = 2 * (2 * (size_t)kATileBytes + 2 * (((size_t)NM * kRowBytes + 1023) / 1024 * 1024)) + 64 + 1024; \
        if ((rc = set_smem_attr(mel_tc_kernel<NM>, smem)) != ADV_OK) return rc;                         \
        mel_tc_kernel<NM><<<grid, kMelThreads, smem, s>>>(a);                                          \
    } while (0)
    if (n_mels <= 64) ADV_MEL(64);
    else if (n_mels <= 80) ADV_MEL(80);
    else ADV_MEL(128);
#undef ADV_MEL
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

int adv_mel_to_channels_last(const float* mel, int batch, int C, int T, int pad, int Cpad, void* out, void* stream) {
    if (!mel || !out || batch <= 0 || C <= 0 || T <= 0 || pad < 0 || Cpad < C) return ADV_ERR_INVALID;
    const long total = (long)batch * (T + 2 * pad) * Cpad;
    int gx = (int)((total + 255) / 256);
    if (gx > 148 * 16) gx = 148 * 16;
    mel_to_cl_kernel<<<gx, 256, 0, (cudaStream_t)stream>>>(mel, batch, C, T, pad, Cpad, (__nv_bfloat16*)out);
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

int adv_avg3_bf16(const void* a, const void* b, const void* c, int64_t n, float act_slope, void* out, void* stream) {
    if (!a || !b || !c || !out || n <= 0 || (n & 1)) return ADV_ERR_INVALID;
    int gx = (int)((n / 2 + 255) / 256);
    if (gx > 148 * 16) gx = 148 * 16;
    avg3_kernel<<<gx, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat162*)a, (const __nv_bfloat162*)b,
                                                     (const __nv_bfloat162*)c, n / 2, act_slope, (__nv_bfloat162*)out);
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

int adv_halo_fix_bf16(void* buf, int batch, int L, int C, int H, int mode, void* stream) {
    if (!buf || batch <= 0 || L <= 0 || C <= 0 || (C & 7) || H <= 0 || (mode != 0 && mode != 1)) return ADV_ERR_INVALID;
    if (mode == 1 && H >= L) return ADV_ERR_SHAPE;   // reflect padding needs pad < length (torch raises too)
    const long total = (long)batch * 2 * H * (C / 8);
    int gx = (int)((total + 255) / 256);
    if (gx > 148 * 8) gx = 148 * 8;
    halo_fix_kernel<<<gx, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)buf, batch, L, C / 8, H, mode);
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

int adv_avg_relayout_bf16(const void* a, const void* b, const void* c, int batch, int L, int C, int h_in, int h_out, int mode,
                          float act_slope, void* out, void* stream) {
    if (!a || !out || batch <= 0 || L <= 0 || C <= 0 || (C & 1) || h_in < 0 || h_out < 0 || (mode != 0 && mode != 1) ||
        (c && !b))
        return ADV_ERR_INVALID;
    if (mode == 1 && h_out >= L) return ADV_ERR_SHAPE;
    const uintptr_t al = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
                         reinterpret_cast<uintptr_t>(out);
    if (C % 8 == 0 && al % 16 == 0) {
        const long total8 = (long)batch * (L + 2 * h_out) * (C / 8);
        int g8 = (int)((total8 + 255) / 256);
        if (g8 > 148 * 16) g8 = 148 * 16;
        avg_relayout8_kernel<<<g8, 256, 0, (cudaStream_t)stream>>>((const int4*)a, (const int4*)b, (const int4*)c, batch, L, C / 8,
                                                                  h_in, h_out, mode, act_slope, (int4*)out);
        ADV_CUDA_CHECK(cudaGetLastError());
        return ADV_OK;
    }
    const long total = (long)batch * (L + 2 * h_out) * (C / 2);
    int gx = (int)((total + 255) / 256);
    if (gx > 148 * 16) gx = 148 * 16;
    avg_relayout_kernel<<<gx, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b,
                                                             (const __nv_bfloat16*)c, batch, L, C / 2, h_in, h_out, mode,
                                                             act_slope, (__nv_bfloat16*)out);
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

int adv_post_conv_tanh(const void* in, const float* w, const float* bias, int batch, int L, int C, int taps, float slope,
                       int pad_reflect, float* out, void* stream) {
    if (!in || !w || !bias || !out || batch <= 0 || L <= 0) return ADV_ERR_INVALID;
    if (C != 32 || taps != 7) return ADV_ERR_UNSUPPORTED;
    const long total = (long)batch * ((L + 3) / 4);
    int gx = (int)((total + 255) / 256);
    if (gx > 148 * 16) gx = 148 * 16;
    post_conv4_kernel<32, 7><<<gx, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)in, w, bias, batch, L, slope,
                                                                  pad_reflect, out);
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

}  // extern "C"
