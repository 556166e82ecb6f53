// Mel front-end in ONE launch: framed STFT (n_fft 1024, any hop / window) -> |X|^p -> filterbank contraction on
// tcgen05 -> log-compress, the spectrum never leaves the SM (audioprocessor.py:38-44 MelSpectrogram;
// hifigan.py:163-178 mel_spectogram).  The two-launch path (adv_stft + adv_mel_project) writes and re-reads
// 8 * F * T bytes per clip (66 MB per 64 x 4 s clips); here the traffic is the waveform in and the mel out.
//
// Structure (one persistent 512-thread CTA per SM, a tile = 64 frame slots):
//   * every warp transforms 2 frame pairs of the tile with the lane-cooperative 1024-point FFT of fft_core.cuh (two real
//     frames per complex transform, 32 values per lane); samples come straight from global memory (coalesced 128-byte
//     loads, reflect padding at the clip edges), the other 15 warps cover the latency;
//   * |X|^p of bins 0..511 is split into bf16 hi + lo parts and written straight into the K-major SWIZZLE_128B operand
//     layout tcgen05.mma consumes: A[64 frames][512 bins] x 2 = 128 KB of shared memory (the Nyquist bin is folded in
//     by the epilogue on CUDA cores, so K stays a multiple of 64).  The FFT's register transposes go THROUGH the two
//     operand rows the pair is about to fill (16 pieces of 256 bytes = one 32 x 32 float plane, addressed with the same
//     128-byte swizzle, conflict-free both ways), so sixteen warps need no scratch of their own - a first version with
//     private scratch could only afford 8 warps and was latency-bound (58 us per 64 clips against 53 us in two launches);
//   * the filterbank sits in shared memory for the CTA's lifetime, split the same way (hi*hi + lo*hi + hi*lo = 16
//     mantissa bits, 1.5e-5 relative: inside the 1e-4 gate) and BAND-COMPRESSED: for each 64-bin K chunk only the mel
//     columns that are non-zero there are stored and multiplied (a triangular bank touches 8 - 32 of its 80 columns per
//     chunk: 36 KB instead of 164 KB, and the MMA N extent shrinks with it).  The host builds the tiles
//     (mel.MelSpectrogram); a dense bank that does not fit is refused (ADV_ERR_UNSUPPORTED) and the caller keeps the
//     two-launch path;
//   * M = 64 per MMA (cta_group::1): accumulator row m lives in TMEM lane (m % 16) + 32 * (m / 16), so warp w reads the
//     rows 16 (w % 4) .. + 15 in its lanes 0..15; the four warps of a TMEM sub-partition split the columns.  The
//     accumulator is zeroed by the epilogue (tcgen05.st) because the chunks write different column ranges;
//   * the MMAs of tile i run while the warps load and transform the first pair of tile i + 1; its epilogue is executed
//     by all sixteen warps right before they overwrite the operand tile.
// Measured and dropped (profiles/r02zr_mel_gen3_variant_slower.jsonl): the same kernel on the generation-3 transform
// (one frame per warp as a 512-point complex transform + real-1024 post pass, 64-bit global loads with the next frame
// prefetched in registers, frames of a tile round-robin over the warps, slot ranges balanced across CTAs) - 53.5 us per
// 64 clips against 45.7 us here, 172 against 147 us per 256: per frame it pays the window multiply, the planar exchanges
// and 17 address computations for the operand stores, which the two-frames-per-transform unit amortises.
#include <cuda_bf16.h>
#include <mutex>
#include <unordered_map>
#include "transform_common.cuh"
#include "umma.cuh"

namespace adv {

using namespace umma;

namespace {

constexpr int kMfThreads = 512, kMfWarps = 16, kMfRows = 64, kMfChunks = 8, kMfPairs = kMfRows / 2 / kMfWarps;
constexpr int kMfChunkBytes = kMfRows * 128;           // one 64-bin chunk of one operand part: 8 KB
constexpr int kMfABytes = kMfChunks * kMfChunkBytes;   // 64 KB per part

struct MelFusedArgs {
    const float* wav;
    int64_t wav_stride;
    int batch, slots_per_clip, total_tiles;
    const unsigned char* fb_tiles;   // dev: per chunk a [N_c][64] bf16 SWIZZLE_128B tile, all hi tiles then all lo tiles
    int fb_bytes, lo_base;           // total bytes (hi + lo), byte offset of the first lo tile
    int chunk_n0[kMfChunks], chunk_n[kMfChunks], chunk_off[kMfChunks];
    const float* fb_nyq;             // dev [NM]: filterbank row of bin 512
    int n_mels, NM;                  // real / padded (multiple of 16) mel count
    float power, clip;
    int log_compress;
    float* out;                      // [B][n_mels][T]
};

__device__ __forceinline__ void tmem_st8_zero(uint32_t taddr) {
    const uint32_t z = 0;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(z) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __noinline__ float mel_pow_generic(float r2, float half_power) { return r2 > 0.f ? powf(r2, half_power) : 0.f; }

// ---- the FFT's two transposes through the operand rows (2p, 2p + 1) of the pair: matrix row k1 (32 floats = 128 bytes)
// lives in piece k1 / 2 (pieces 0..7: the eight chunks of the hi part, 8..15: of the lo part; the pair's two rows are
// adjacent there), at 128 * (k1 % 2), its 16-byte groups XOR-swizzled with k1 % 8.  `base` = A_hi + 128 * (2p).
__device__ __forceinline__ void arow_store_cols(const float2* v, int l, unsigned char* base, bool imag) {
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) {
        unsigned char* p = base + ((k1 >> 1) & 7) * kMfChunkBytes + (k1 >> 4) * kMfABytes + (k1 & 1) * 128 +
                           ((((l >> 2) ^ (k1 & 7)) << 4) | ((l & 3) << 2));
        *reinterpret_cast<float*>(p) = imag ? v[k1].y : v[k1].x;
    }
}
__device__ __forceinline__ void arow_load_rows(float2* v, int l, const unsigned char* base, bool imag) {
    const unsigned char* rb = base + ((l >> 1) & 7) * kMfChunkBytes + (l >> 4) * kMfABytes + (l & 1) * 128;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 q = *reinterpret_cast<const float4*>(rb + ((i ^ (l & 7)) << 4));
        float2* d = v + 4 * i;
        if (imag) { d[0].y = q.x; d[1].y = q.y; d[2].y = q.z; d[3].y = q.w; }
        else      { d[0].x = q.x; d[1].x = q.y; d[2].x = q.z; d[3].x = q.w; }
    }
}

__global__ void __launch_bounds__(kMfThreads, 1)
mel_fused_kernel(PlanDev P, MelFusedArgs a) {
    constexpr int NF = 1024;
    using G = Geo<NF>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* A_hi = smem;
    unsigned char* A_lo = smem + kMfABytes;
    unsigned char* FB = smem + 2 * kMfABytes;
    Carver cv{FB + a.fb_bytes};
    float2* tw_s = cv.take<float2>(32 * G::LANES);
    float* win_s = cv.take<float>(NF);
    float* nyq_s = cv.take<float>(2 * kMfRows);        // |X[512]|^p of the tile's rows, two tiles deep
    float* fbn_s = cv.take<float>(a.NM);
    uint64_t* bar = cv.take<uint64_t>(1);
    uint32_t* tmem_slot = cv.take<uint32_t>(1);

    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        bar_init(bar, 1);
        bar_init_fence();
    }
    for (int i = tid; i < 32 * G::LANES / 2; i += kMfThreads) cp_async16(tw_s + 2 * i, P.tw + 2 * i);
    for (int i = tid; i < NF / 4; i += kMfThreads) cp_async16(win_s + 4 * i, P.window + 4 * i);
    for (int i = tid; i < a.fb_bytes / 16; i += kMfThreads) cp_async16(FB + 16 * i, a.fb_tiles + 16 * (size_t)i);
    for (int i = tid; i < a.NM; i += kMfThreads) fbn_s[i] = a.fb_nyq[i];
    if (w == 0) tmem_alloc(tmem_slot, 128);
    cp_async_wait_all();
    umma::fence_async_smem();   // the filterbank tiles are read by the async proxy
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_d = *tmem_slot;
    const int sub = w & 3, quarter = w >> 2;                        // TMEM sub-partition, column quarter of this warp
    const int c_lo = ((a.NM / 8) * quarter / 4) * 8, c_hi = ((a.NM / 8) * (quarter + 1) / 4) * 8;
    const uint32_t trow = tmem_d + ((uint32_t)(32 * sub) << 16);
    for (int c0 = c_lo; c0 < c_hi; c0 += 8) tmem_st8_zero(trow + c0);
    tmem_st_wait();

    const TwSmem<G::LANES> tw{tw_s, lane};
    const int hop = P.hop, n_in = P.n_in, T = P.T, spc = a.slots_per_clip;
    const bool p2 = a.power == 2.0f, p1 = a.power == 1.0f;
    const float half_power = 0.5f * a.power;

    // windowed samples of the pair (tile, q) of this warp: .x = frame t0, .y = frame t0 + 1; zeros outside the work list
    auto load_pair = [&](int tile, int q, float2 (&v)[32]) {
        const long slot = (long)tile * kMfRows + 2 * (kMfPairs * w + q);
        const int b = (int)(slot / spc), t0 = (int)(slot - (long)b * spc);
        if (!(b < a.batch && t0 < T)) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = make_float2(0.f, 0.f);
            return;
        }
        const float* row = a.wav + (size_t)b * a.wav_stride;
        const int base = t0 * hop - NF / 2;
        if (base >= 0 && base + hop + NF <= n_in && t0 + 1 < T) {
            const float* pa = row + base + lane;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i].x = __ldg(pa + 32 * i);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i].y = __ldg(pa + hop + 32 * i);
        } else {
            const bool has_b = t0 + 1 < T;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                int ia = base + lane + 32 * i, ib = ia + hop;
                ia = ia < 0 ? -ia : (ia >= n_in ? 2 * (n_in - 1) - ia : ia);
                ib = ib < 0 ? -ib : (ib >= n_in ? 2 * (n_in - 1) - ib : ib);
                v[i].x = (ia >= 0 && ia < n_in) ? __ldg(row + ia) : 0.f;
                v[i].y = (has_b && ib >= 0 && ib < n_in) ? __ldg(row + ib) : 0.f;
            }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float ww = win_s[lane + 32 * i];
            v[i].x *= ww;
            v[i].y *= ww;
        }
    };
    // epilogue of the tile whose MMAs were committed last: TMEM -> (+ Nyquist term) -> log -> out, then zero the columns
    auto epilogue = [&](int tile, int par) {
        const int r = 16 * sub + lane;                       // accumulator row of this lane (lanes 0..15)
        const long slot = (long)tile * kMfRows + r;
        const int b = (int)(slot / spc), t = (int)(slot - (long)b * spc);
        const bool live = lane < 16 && b < a.batch && t < T;
        const float pn = nyq_s[par * kMfRows + (r & (kMfRows - 1))];
        fence_after_sync();
        for (int c0 = c_lo; c0 < c_hi; c0 += 8) {
            float v[8];
            tmem_ld8(trow + c0, v);
            tmem_st8_zero(trow + c0);
            if (live) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int mel = c0 + j;
                    if (mel < a.n_mels) {
                        float y = fmaf(pn, fbn_s[mel], v[j]);
                        if (a.log_compress) y = logf(fmaxf(y, a.clip));
                        a.out[((size_t)b * a.n_mels + mel) * T + t] = y;
                    }
                }
            }
        }
        tmem_st_wait();
        fence_before_sync();
    };

    int tile = blockIdx.x;
    int it = 0, prev_tile = -1;
    for (; tile < a.total_tiles; tile += gridDim.x, ++it) {
        const int par = it & 1;
#pragma unroll 1
        for (int q = 0; q < kMfPairs; ++q) {
            float2 v[32];
            load_pair(tile, q, v);
            fwd_cols<NF>(v, tw);
            if (q == 0 && prev_tile >= 0) {   // the operand tile is still being read by the previous tile's MMAs
                bar_wait(bar, (it - 1) & 1);
                epilogue(prev_tile, par ^ 1);
            }
            const int rowa = 2 * (kMfPairs * w + q), rowb = rowa + 1;
            unsigned char* arow = A_hi + rowa * 128;
            __syncwarp();
            arow_store_cols(v, lane, arow, false);
            __syncwarp();
            arow_load_rows(v, lane, arow, false);
            __syncwarp();
            arow_store_cols(v, lane, arow, true);
            __syncwarp();
            arow_load_rows(v, lane, arow, true);
            fwd_rows<NF>(v);
            float2 xa[17], xb[17];
            split_regs<NF>(v, lane, xa, xb);
            __syncwarp();   // every lane has read its transposed rows: the pair's operand rows may be overwritten
            // byte offsets of this lane's element inside a chunk, for even / odd slots (bins l + 64 c and l + 32 + 64 c)
            const int j0 = lane >> 3, e2 = (lane & 7) * 2;
            const int oa0 = rowa * 128 + ((j0 ^ (rowa & 7)) << 4) + e2, oa1 = rowa * 128 + (((j0 + 4) ^ (rowa & 7)) << 4) + e2;
            const int ob0 = rowb * 128 + ((j0 ^ (rowb & 7)) << 4) + e2, ob1 = rowb * 128 + (((j0 + 4) ^ (rowb & 7)) << 4) + e2;
#pragma unroll
            for (int i = 0; i < 17; ++i) {
                const float r2a = fmaf(xa[i].x, xa[i].x, xa[i].y * xa[i].y), r2b = fmaf(xb[i].x, xb[i].x, xb[i].y * xb[i].y);
                float pa, pb;
                if (p2) { pa = r2a; pb = r2b; }
                else if (p1) { pa = r2a * rsqrt_ftz(fmaxf(r2a, 1e-37f)); pb = r2b * rsqrt_ftz(fmaxf(r2b, 1e-37f)); }
                else { pa = mel_pow_generic(r2a, half_power); pb = mel_pow_generic(r2b, half_power); }
                if (i < 16) {
                    const __nv_bfloat16 ha = __float2bfloat16_rn(pa), hb = __float2bfloat16_rn(pb);
                    const __nv_bfloat16 la = __float2bfloat16_rn(pa - __bfloat162float(ha));
                    const __nv_bfloat16 lb = __float2bfloat16_rn(pb - __bfloat162float(hb));
                    const int cb = (i >> 1) * kMfChunkBytes;
                    const int oa = cb + ((i & 1) ? oa1 : oa0), ob = cb + ((i & 1) ? ob1 : ob0);
                    *reinterpret_cast<__nv_bfloat16*>(A_hi + oa) = ha;
                    *reinterpret_cast<__nv_bfloat16*>(A_lo + oa) = la;
                    *reinterpret_cast<__nv_bfloat16*>(A_hi + ob) = hb;
                    *reinterpret_cast<__nv_bfloat16*>(A_lo + ob) = lb;
                } else if (lane == 0) {   // bin 512
                    nyq_s[par * kMfRows + rowa] = pa;
                    nyq_s[par * kMfRows + rowb] = pb;
                }
            }
        }
        umma::fence_async_smem();
        fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            fence_after_sync();
#pragma unroll 1
            for (int c = 0; c < kMfChunks; ++c) {
                const int n = a.chunk_n[c];
                if (n <= 0) continue;
                const uint32_t idesc = make_idesc(FMT_BF16, kMfRows, (uint32_t)n);
                const uint64_t dah = make_desc_sw128(smem_addr(A_hi + c * kMfChunkBytes));
                const uint64_t dal = make_desc_sw128(smem_addr(A_lo + c * kMfChunkBytes));
                const uint64_t dbh = make_desc_sw128(smem_addr(FB + a.chunk_off[c]));
                const uint64_t dbl = make_desc_sw128(smem_addr(FB + a.lo_base + a.chunk_off[c]));
                const uint32_t d = tmem_d + (uint32_t)a.chunk_n0[c];
#pragma unroll
                for (int k = 0; k < 4; ++k) {   // K = 16 bf16 = 32 bytes per MMA
                    mma_f16(d, dal + 2 * k, dbh + 2 * k, idesc, 1);   // small terms first
                    mma_f16(d, dah + 2 * k, dbl + 2 * k, idesc, 1);
                    mma_f16(d, dah + 2 * k, dbh + 2 * k, idesc, 1);
                }
            }
            mma_commit(bar);
        }
        prev_tile = tile;
    }
    if (prev_tile >= 0) {
        bar_wait(bar, (it - 1) & 1);
        epilogue(prev_tile, (it - 1) & 1);
    }
    __syncthreads();
    if (w == 0) tmem_dealloc(tmem_d, 128);
}

}  // namespace

}  // namespace adv

using namespace adv;

extern "C" int adv_mel_fused(const adv_plan* p, const float* wav, int64_t wav_stride, int batch, const void* fb_tiles,
                             int fb_bytes, int lo_base, const int* chunk_table, const float* fb_nyq, int n_mels,
                             float power, int log_compress, float clip, float* out, void* stream) {
    if (!p || !wav || !fb_tiles || !chunk_table || !fb_nyq || !out || batch <= 0 || n_mels <= 0) return ADV_ERR_INVALID;
    if (p->d.n_fft != 1024 || p->d.n_in <= 0) return ADV_ERR_UNSUPPORTED;
    if (n_mels > 128 || fb_bytes <= 0 || fb_bytes % 1024 != 0 || lo_base <= 0 || lo_base % 1024 != 0) return ADV_ERR_SHAPE;
    MelFusedArgs a;
    a.wav = wav; a.wav_stride = wav_stride; a.batch = batch;
    a.slots_per_clip = (p->d.T + 1) & ~1;
    const long slots = (long)a.slots_per_clip * batch;
    const long tiles = (slots + kMfRows - 1) / kMfRows;
    if (tiles > 0x3fffffffL) return ADV_ERR_UNSUPPORTED;
    a.total_tiles = (int)tiles;
    a.fb_tiles = (const unsigned char*)fb_tiles; a.fb_bytes = fb_bytes; a.lo_base = lo_base;
    a.NM = (n_mels + 15) & ~15;
    for (int c = 0; c < kMfChunks; ++c) {
        a.chunk_n0[c] = chunk_table[3 * c]; a.chunk_n[c] = chunk_table[3 * c + 1]; a.chunk_off[c] = chunk_table[3 * c + 2];
        if (a.chunk_n[c] < 0 || a.chunk_n[c] % 8 != 0 || a.chunk_n0[c] < 0 || a.chunk_n0[c] + a.chunk_n[c] > a.NM ||
            a.chunk_off[c] % 1024 != 0 || a.chunk_off[c] + a.chunk_n[c] * 128 > lo_base || 2 * lo_base > fb_bytes)
            return ADV_ERR_SHAPE;
    }
    a.fb_nyq = fb_nyq; a.n_mels = n_mels; a.power = power; a.clip = clip; a.log_compress = log_compress; a.out = out;
    const size_t smem = 1024 + 2 * (size_t)kMfABytes + fb_bytes + al16(8 * 32 * 32) + al16(4 * 1024) +
                        al16(4 * 2 * kMfRows) + al16(4 * a.NM) + 16 + 16;
    if (smem > 227 * 1024) return ADV_ERR_UNSUPPORTED;   // dense filterbank: the caller keeps the two-launch path
    int rc = set_smem(mel_fused_kernel, smem);
    if (rc != ADV_OK) return rc;
    const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
    mel_fused_kernel<<<grid, kMfThreads, smem, (cudaStream_t)stream>>>(p->d, a);
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}
