// Warp-specialised, TMA-fed, persistent tcgen05 implicit-GEMM conv1d (bf16, channels-last) for sm_100a.
//
// This is the production path of the HiFi-GAN generator's convolutions (the first-generation kernel in
// gemm_kernels.cu gathers its operands with ordinary loads and remains the fallback for reflect padding
// and odd channel counts).  Reference call site: hifi_gan.decode_batch, hifigan.py:180.
//
//   D[128 positions][BN channels] (+)= A[128][taps*Cin] * W[BN][taps*Cin]^T        fp32 in TMEM
//
// * A comes straight from the channels-last activation [B][L][Cin] through a 3-D tensor map
//   (box = BK channels x 128 positions x 1 clip): a tap is just a shifted row coordinate, and rows
//   before 0 / past L are zero-filled by the TMA unit = the conv's zero padding, per clip.
// * W [N][taps*Cin] through a 2-D tensor map (box = BK x BN).  Both land in the canonical K-major
//   128-byte (Cin >= 64) or 64-byte (Cin = 32) swizzled layout that tcgen05.mma reads.
// * Roles: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-5 = epilogue
//   (TMEM -> registers -> +bias, +residual -> raw bf16 and / or LeakyReLU'd bf16 stores).
//   STAGES-deep full/empty mbarrier ring between producer and MMA; two TMEM accumulator buffers so
//   the epilogue of tile i overlaps the MMAs of tile i+1; CTAs are persistent over a static tile list.
// * LeakyReLU is applied by the PRODUCING layer's epilogue (it can write the activated tensor next
//   to / instead of the raw one), so consumers load pure TMA tiles.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include "adv_internal.cuh"
#include "umma.cuh"

namespace adv {

using namespace umma;

constexpr int kConvThreads = 192;  // 6 warps: producer, mma, 4 x epilogue

struct ConvTmaArgs {
    const float* bias;            // [N] or null
    const __nv_bfloat16* resid;   // [B][L][N] or null
    __nv_bfloat16* out_raw;       // [B][L][N] or null
    __nv_bfloat16* out_act;       // [B][L][N] or null: LeakyReLU(act_slope) of the result
    int B, L, Cin, taps, dil, center, N;
    int tiles_l, tiles_n;         // tiles per clip along L, tiles along N
    float act_slope, out_scale;
};

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_addr(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_addr(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_addr(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_addr(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
// K-major descriptor for 128-byte (BK = 64) or 64-byte (BK = 32) swizzled rows
template <int BK>
__device__ __forceinline__ uint64_t make_desc_k(uint32_t saddr) {
    constexpr uint32_t row_bytes = BK * 2;
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((8 * row_bytes) >> 4) << 32;         // stride between 8-row groups
    d |= (uint64_t)1 << 46;                              // Blackwell descriptor version
    d |= (uint64_t)(row_bytes == 128 ? 2 : 4) << 61;     // SWIZZLE_128B / SWIZZLE_64B
    return d;
}

// Epilogue of one 128 x BN tile for one warp (TMEM lanes 32*quad ..): +bias, +residual, *scale, raw and / or
// LeakyReLU'd bf16 stores.  The residual row is fetched before the accumulator is waited for.
struct EpiArgs {
    const float* bias;
    const __nv_bfloat16* resid;
    __nv_bfloat16* out_raw;
    __nv_bfloat16* out_act;
    int L, N;
    float act_slope, out_scale;
};
template <int BN>
__device__ __forceinline__ void conv_epilogue(const EpiArgs& a, int b, int l0, int tn, uint32_t tmem_acc, int quad,
                                              int lane, uint64_t* tfull_bar, uint32_t parity) {
    const int l = l0 + quad * 32 + lane;
    const bool row_ok = l < a.L;
    const size_t rowoff = ((size_t)b * a.L + l) * a.N + (size_t)tn * BN;
    const bool has_res = a.resid != nullptr && row_ok;
    // the residual row does not depend on the MMAs: fetch its first 32 columns while they finish
    int4 rnext[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
        rnext[j] = has_res ? __ldg(reinterpret_cast<const int4*>(a.resid + rowoff) + j) : make_int4(0, 0, 0, 0);
    bar_wait(tfull_bar, parity);
    fence_after_sync();
    const uint32_t trow = tmem_acc + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
        int4 rcur[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) rcur[j] = rnext[j];
        if (c0 + 32 < BN) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                rnext[j] = has_res ? __ldg(reinterpret_cast<const int4*>(a.resid + rowoff + c0 + 32) + j)
                                   : make_int4(0, 0, 0, 0);
        }
        float v[32];
        tmem_ld32(trow + c0, v);
        if (row_ok) {
            const int n = tn * BN + c0;
            if (a.bias != nullptr) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias + n) + j);
                    v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
                }
            }
            if (has_res) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const __nv_bfloat162* rp = reinterpret_cast<const __nv_bfloat162*>(&rcur[q]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = __bfloat1622float2(rp[j]);
                        v[8 * q + 2 * j] += f.x;
                        v[8 * q + 2 * j + 1] += f.y;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= a.out_scale;
            if (a.out_raw != nullptr) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    int4 o;
                    __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                    for (int j = 0; j < 4; ++j) op[j] = __floats2bfloat162_rn(v[8 * q + 2 * j], v[8 * q + 2 * j + 1]);
                    *(reinterpret_cast<int4*>(a.out_raw + rowoff + c0) + q) = o;
                }
            }
            if (a.out_act != nullptr) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    int4 o;
                    __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float x0 = v[8 * q + 2 * j], x1 = v[8 * q + 2 * j + 1];
                        op[j] = __floats2bfloat162_rn(x0 > 0.f ? x0 : x0 * a.act_slope, x1 > 0.f ? x1 : x1 * a.act_slope);
                    }
                    *(reinterpret_cast<int4*>(a.out_act + rowoff + c0) + q) = o;
                }
            }
        }
    }
}

// ---- epilogue with TMA stores -------------------------------------------------------------------------------------
// What ncu says bounds the thread-per-row epilogue (gpurun_out/r02zd, profiles/r02zd_*): L1 is the busiest unit of the
// kernel (66 - 72 %), every 16-byte store lands in a different 128-byte line (31.6 sectors per request), DRAM sits at
// 39 %.  Variants that re-read a staged tile with the LSU lost on instruction count.  Here the warp writes its 32 x 32
// output block (raw and / or activated) into a 2 KB SWIZZLE_64B tile and ONE lane hands it to the TMA engine
// (cp.async.bulk.tensor shared -> global, 3-D map [C][L][B], box 32 x 32 x 1): four STS.128 per thread replace four
// uncoalesced STG.128, rows past the end of the clip are clipped by the tensor map, and the store runs asynchronously under
// the next 32 columns' TMEM load and arithmetic.  `stg`: 4 KB per warp (raw tile, activated tile).
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_addr(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Residual block through the TMA engine too (res_tma): the warp's 32 x 32 block of the residual tensor is requested one
// block ahead into one of two 2 KB tiles (mbarrier per tile, `rcount` = blocks requested so far by this warp: tile
// rcount & 1, parity (rcount >> 1) & 1) and read back with one conflict-free LDS.128 per 8 columns - the per-thread
// LDG.128 form touches 11 sectors per request.  `stg`: 8 KB per warp = raw tile, activated tile, 2 residual tiles;
// `rbar`: this warp's two mbarriers.
struct EpiTma {
    const CUtensorMap* map_raw;
    const CUtensorMap* map_act;
    const CUtensorMap* map_res;
    int res_tma;
};
template <int BN, bool RES_TMA>
__device__ __forceinline__ void conv_epilogue_tma(const EpiArgs& a, const EpiTma& m, int b, int l0, int tn, uint32_t tmem_acc,
                                                  int quad, int lane, uint64_t* tfull_bar, uint32_t parity, unsigned char* stg,
                                                  uint64_t* rbar, uint32_t& rcount) {
    const int lrow = l0 + quad * 32;
    const int l = lrow + lane;
    const bool row_ok = l < a.L;
    const size_t rowoff = ((size_t)b * a.L + l) * a.N + (size_t)tn * BN;
    const bool has_res = a.resid != nullptr;
    constexpr bool res_tma = RES_TMA;                    // (the caller instantiates it only for layers with a residual)
    const bool res_ldg = has_res && !RES_TMA && row_ok;
    auto request_res = [&](uint32_t k, int c0) {   // block k of this warp: columns tn * BN + c0 .. + 31
        if (lane == 0) {
            bar_expect_tx(&rbar[k & 1], 2048);
            tma_load_3d(stg + 4096 + (k & 1) * 2048, m.map_res, &rbar[k & 1], tn * BN + c0, lrow, b);
        }
    };
    int4 rnext[4];
    if (res_tma) request_res(rcount, 0);
#pragma unroll
    for (int j = 0; j < 4; ++j)
        rnext[j] = res_ldg ? __ldg(reinterpret_cast<const int4*>(a.resid + rowoff) + j) : make_int4(0, 0, 0, 0);
    bar_wait(tfull_bar, parity);
    fence_after_sync();
    const uint32_t trow = tmem_acc + ((uint32_t)(quad * 32) << 16);
    unsigned char* tile_raw = stg + lane * 64;
    unsigned char* tile_act = stg + 2048 + lane * 64;
    const int sw = (lane >> 1) & 3;   // SWIZZLE_64B: chunk ^= bits [7, 9) of the byte address = (row >> 1) & 3
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
        int4 rcur[4];
        if (res_tma) {
            if (c0 + 32 < BN) request_res(rcount + 1, c0 + 32);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) rcur[j] = rnext[j];
            if (c0 + 32 < BN) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    rnext[j] = res_ldg ? __ldg(reinterpret_cast<const int4*>(a.resid + rowoff + c0 + 32) + j)
                                       : make_int4(0, 0, 0, 0);
            }
        }
        float v[32];
        tmem_ld32(trow + c0, v);
        const int n = tn * BN + c0;
        if (a.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias + n) + j);
                v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
            }
        }
        if (res_tma) {
            bar_wait(&rbar[rcount & 1], (rcount >> 1) & 1);
            const unsigned char* rt = stg + 4096 + (rcount & 1) * 2048 + lane * 64;
#pragma unroll
            for (int q = 0; q < 4; ++q) rcur[q] = *reinterpret_cast<const int4*>(rt + ((q ^ sw) << 4));
            ++rcount;
        }
        if (has_res) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const __nv_bfloat162* rp = reinterpret_cast<const __nv_bfloat162*>(&rcur[q]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __bfloat1622float2(rp[j]);
                    v[8 * q + 2 * j] += f.x;
                    v[8 * q + 2 * j + 1] += f.y;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= a.out_scale;
        // the previous block's stores have read the tiles
        if (lane == 0) bulk_wait_read0();
        __syncwarp();   // (also: every lane has read its residual row before that tile is requested again)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int4 o, oa;
            __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(&o);
            __nv_bfloat162* oq = reinterpret_cast<__nv_bfloat162*>(&oa);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float x0 = v[8 * q + 2 * j], x1 = v[8 * q + 2 * j + 1];
                op[j] = __floats2bfloat162_rn(x0, x1);
                oq[j] = __floats2bfloat162_rn(x0 > 0.f ? x0 : x0 * a.act_slope, x1 > 0.f ? x1 : x1 * a.act_slope);
            }
            if (a.out_raw != nullptr) *reinterpret_cast<int4*>(tile_raw + ((q ^ sw) << 4)) = o;
            if (a.out_act != nullptr) *reinterpret_cast<int4*>(tile_act + ((q ^ sw) << 4)) = oa;
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            if (a.out_raw != nullptr) tma_store_3d(m.map_raw, stg, n, lrow, b);
            if (a.out_act != nullptr) tma_store_3d(m.map_act, stg + 2048, n, lrow, b);
            bulk_commit();
        }
    }
}

// ---- epilogue with the residual row held in registers --------------------------------------------------------------
// The layers that add a residual and write two outputs are HBM-bound (4.3 GB per launch at 256 clips), and the epilogue
// above moves 3/4 of those bytes with 64 bytes per thread in flight and one exposed memory latency per 32 columns: 39 %
// of the DRAM throughput, 14 % tensor activity on the 128-channel convs (profiles/r02zb_vocoder_launches_b256.csv).  Here a
// thread's WHOLE residual row (BN / 8 int4) is requested one tile ahead - while the previous tile is converted and stored -
// so 128 threads keep 16 - 64 KB in flight per CTA and no load is waited for.  One CTA per SM: the registers are there.
// NH = 1: the warp owns all BN columns of its 32 rows; NH = 2: two warps share a TMEM lane quarter, `half` selects the
// BN / 2 columns of this one.
template <int BN, int NH = 1>
__device__ __forceinline__ void epi_load_resid(const EpiArgs& a, int b, int l0, int tn, int quad, int lane, bool live,
                                               int4 (&r)[BN / 8 / NH], int half = 0) {
    const int l = l0 + quad * 32 + lane;
    const bool ok = live && a.resid != nullptr && l < a.L;
    const int4* p = reinterpret_cast<const int4*>(a.resid + ((size_t)b * a.L + (ok ? l : 0)) * a.N + (size_t)tn * BN +
                                                  half * (BN / NH));
#pragma unroll
    for (int j = 0; j < BN / 8 / NH; ++j) r[j] = ok ? __ldg(p + j) : make_int4(0, 0, 0, 0);
}
// Coalesced form: ncu on the residual convs (gpurun_out/r02zd) shows what bounds them - every 16-byte store of the
// thread-per-row epilogue lands in a different 128-byte line (31.6 sectors per request), the LSU data pipe is 61 % busy and
// L1 is the busiest unit of the kernel (66 - 72 %) while DRAM sits at 39 %.  Here a warp stages 32 output columns of its 32
// rows in a private 2 KB shared-memory tile (16-byte chunks XOR-swizzled with the row) and writes it out with 4 lanes per
// row: a store instruction covers 8 rows x 64 contiguous bytes = 16 full sectors instead of 32 half-used ones in 32 lines.
// `stg` = this warp's tile.
constexpr int kEpiStageBytes = 2048;
template <int BN, int NH = 1>
__device__ __forceinline__ void conv_epilogue_pre(const EpiArgs& a, int b, int l0, int tn, uint32_t tmem_acc, int quad,
                                                  int lane, uint64_t* tfull_bar, uint32_t parity,
                                                  const int4 (&res)[BN / 8 / NH], unsigned char* stg, int half = 0) {
    constexpr int W = 32;                         // columns staged per pass (one tcgen05.ld.x32)
    constexpr int CH = W / 8;                     // 16-byte chunks per staged row
    constexpr int RPI = 32 / CH;                  // rows written per store instruction
    const int lrow = l0 + quad * 32;              // first row of this warp
    const bool has_res = a.resid != nullptr;
    bar_wait(tfull_bar, parity);
    fence_after_sync();
    const uint32_t trow = tmem_acc + ((uint32_t)(quad * 32) << 16);
    const int srow = lane / CH, sch = lane % CH;  // this lane's (row offset, chunk) in the write-out
    const int cbase = half * (BN / NH);           // first column of this warp
#pragma unroll
    for (int pp = 0; pp < BN / NH; pp += W) {
        const int p0 = cbase + pp;
        int4 act[CH];
#pragma unroll
        for (int h = 0; h < W / 32; ++h) {
            const int c0 = p0 + 32 * h;
            float v[32];
            tmem_ld32(trow + c0, v);
            const int n = tn * BN + c0;
            if (a.bias != nullptr) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias + n) + j);
                    v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
                }
            }
            if (has_res) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const __nv_bfloat162* rp = reinterpret_cast<const __nv_bfloat162*>(&res[(pp + 32 * h) / 8 + q]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = __bfloat1622float2(rp[j]);
                        v[8 * q + 2 * j] += f.x;
                        v[8 * q + 2 * j + 1] += f.y;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= a.out_scale;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int4 o, oa;
                __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(&o);
                __nv_bfloat162* oq = reinterpret_cast<__nv_bfloat162*>(&oa);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float x0 = v[8 * q + 2 * j], x1 = v[8 * q + 2 * j + 1];
                    op[j] = __floats2bfloat162_rn(x0, x1);
                    oq[j] = __floats2bfloat162_rn(x0 > 0.f ? x0 : x0 * a.act_slope, x1 > 0.f ? x1 : x1 * a.act_slope);
                }
                const int ch = 4 * h + q;
                act[ch] = oa;
                if (a.out_raw != nullptr) *reinterpret_cast<int4*>(stg + lane * (CH * 16) + ((ch ^ (lane & (CH - 1))) << 4)) = o;
            }
        }
        // write-out: raw tile, then (through the same buffer) the activated tile
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            __nv_bfloat16* dst = pass == 0 ? a.out_raw : a.out_act;
            if (dst == nullptr) continue;   // (uniform)
            if (pass == 1) {
                __syncwarp();   // the raw tile has been read out
#pragma unroll
                for (int ch = 0; ch < CH; ++ch)
                    *reinterpret_cast<int4*>(stg + lane * (CH * 16) + ((ch ^ (lane & (CH - 1))) << 4)) = act[ch];
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 32 / RPI; ++i) {
                const int r = i * RPI + srow;
                const int l = lrow + r;
                const int4 o = *reinterpret_cast<const int4*>(stg + r * (CH * 16) + ((sch ^ (r & (CH - 1))) << 4));
                if (l < a.L)
                    *reinterpret_cast<int4*>(dst + ((size_t)b * a.L + l) * a.N + (size_t)tn * BN + p0 + sch * 8) = o;
            }
        }
        __syncwarp();   // the tile is free for the next pass
    }
}

template <int BN, int BK, int STAGES>
__global__ void __launch_bounds__(kConvThreads, 1)
conv1d_tma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, ConvTmaArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kABytes = 128 * BK * 2, kBBytes = BN * BK * 2, kStage = kABytes + kBBytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * kStage);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    constexpr uint32_t kCols = 2 * BN < 32 ? 32 : 2 * BN;  // two accumulator buffers

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            bar_init(&full[s], 1);
            bar_init(&empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            bar_init(&tfull[i], 1);
            bar_init(&tempty[i], 4);
        }
        bar_init_fence();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kCols);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int cblocks = a.Cin / BK;
    const int nkb = a.taps * cblocks;
    const long total_tiles = (long)a.B * a.tiles_l * a.tiles_n;

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0) {
            uint32_t it = 0;
            for (long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int tn = (int)(tile % a.tiles_n);
                const long tl = tile / a.tiles_n;
                const int b = (int)(tl / a.tiles_l), l0 = (int)(tl % a.tiles_l) * 128;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES, ph = (it / STAGES) & 1;
                    bar_wait(&empty[s], ph ^ 1);
                    bar_expect_tx(&full[s], kStage);
                    const int tap = kb / cblocks, cb = kb - tap * cblocks;
                    unsigned char* sA = smem + s * kStage;
                    tma_load_3d(sA, &map_a, &full[s], cb * BK, l0 + (tap - a.center) * a.dil, b);
                    tma_load_2d(sA + kABytes, &map_w, &full[s], kb * BK, tn * BN);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer ------------------------------
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(FMT_BF16, 128, BN);
            uint32_t it = 0, tcount = 0;
            for (long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
                const uint32_t ab = tcount & 1, aph = (tcount >> 1) & 1;
                bar_wait(&tempty[ab], aph ^ 1);  // the epilogue has drained this accumulator buffer
                fence_after_sync();
                const uint32_t tmem_d = tmem_base + ab * BN;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES, ph = (it / STAGES) & 1;
                    bar_wait(&full[s], ph);
                    fence_after_sync();
                    const uint32_t sA = smem_addr(smem + s * kStage);
                    const uint64_t da = make_desc_k<BK>(sA), db = make_desc_k<BK>(sA + kABytes);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        mma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    mma_commit(&empty[s]);  // frees the stage once these MMAs have read it
                }
                mma_commit(&tfull[ab]);     // accumulator complete
            }
        }
    } else {
        // ------------------------------ epilogue (4 warps) ------------------------------
        const int quad = warp & 3;  // TMEM lanes 32*quad .. 32*quad+31 are the ones this warp may read
        uint32_t tcount = 0;
        for (long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
            const int tn = (int)(tile % a.tiles_n);
            const long tl = tile / a.tiles_n;
            const int b = (int)(tl / a.tiles_l), l0 = (int)(tl % a.tiles_l) * 128;
            const uint32_t ab = tcount & 1, aph = (tcount >> 1) & 1;
            const EpiArgs ea{a.bias, a.resid, a.out_raw, a.out_act, a.L, a.N, a.act_slope, a.out_scale};
            conv_epilogue<BN>(ea, b, l0, tn, tmem_base + ab * BN, quad, lane, &tfull[ab], aph);
            fence_before_sync();
            __syncwarp();
            if (lane == 0) bar_arrive(&tempty[ab]);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kCols);
}

// ---------------------------------------------------------------------------------------------------
// Narrow-channel variant (C_in = C_out = C in {32, 64}: the MRF stages that are memory-, not tensor-bound).
// The general kernel re-fetches a shifted 128-row A tile per tap (up to 11x the activation through L2) and
// the weight tile per K-block.  Here a persistent CTA keeps ALL weights resident in shared memory and
// fetches, per tile, ONE slab of 128 + 2*halo activation rows; tap j is the same slab read through a
// descriptor whose start address is advanced by j*dil rows (the swizzle phase follows the address).
// ---------------------------------------------------------------------------------------------------
struct SlabArgs {
    EpiArgs e;
    int B, taps, dil, halo, rows;  // rows = 128 + 2*halo (TMA box height)
    int tiles_l;
    int epi_tma;   // outputs through TMA stores (conv_epilogue_tma)
    int res_tma;   // ... and the residual block through TMA loads
};

// Row-shifted start addresses need nothing special in the descriptor: measured on B200, the swizzle XOR is
// taken from the absolute shared-memory address bits (base_offset stays 0; setting it to (addr >> 7) & 7
// breaks the result), for both the 128-byte and the 64-byte modes and for odd row shifts.

template <int C, int STAGES>
__global__ void __launch_bounds__(kConvThreads, 2)   // two CTAs per SM where shared memory allows: <= 168 registers
conv1d_slab_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                   const __grid_constant__ CUtensorMap map_or, const __grid_constant__ CUtensorMap map_oa,
                   const __grid_constant__ CUtensorMap map_rs, SlabArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kRow = C * 2, kWTile = C * kRow;
    const int slab_stride = (a.rows * kRow + 1023) & ~1023;
    unsigned char* wsm = smem;                               // [taps][C x C] weights, resident
    unsigned char* slabs = smem + a.taps * kWTile;           // [STAGES][slab_stride]
    uint64_t* full = reinterpret_cast<uint64_t*>(slabs + STAGES * slab_stride);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint64_t* wbar = tempty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
    constexpr uint32_t kCols = 2 * C;
    // per-warp 8 KB of epilogue tiles (conv_epilogue_tma: raw, activated, 2 residual tiles; conv_epilogue_pre: the first 2 KB)
    // and the residual tiles' mbarriers
    uint64_t* rbars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));
    unsigned char* stg = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(rbars + 8) + 1023) & ~uintptr_t(1023)) +
                         ((threadIdx.x >> 5) & 3) * 8192;
    uint64_t* rbar = rbars + 2 * ((threadIdx.x >> 5) & 3);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            bar_init(&full[s], 1);
            bar_init(&empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            bar_init(&tfull[i], 1);
            bar_init(&tempty[i], 4);
        }
        bar_init(wbar, 1);
        for (int i = 0; i < 8; ++i) bar_init(&rbars[i], 1);
        bar_init_fence();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kCols);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const long total_tiles = (long)a.B * a.tiles_l;

    if (warp == 0) {
        if (lane == 0) {
            bar_expect_tx(wbar, a.taps * kWTile);
            for (int tap = 0; tap < a.taps; ++tap) tma_load_2d(wsm + tap * kWTile, &map_w, wbar, tap * C, 0);
            uint32_t it = 0;
            for (long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int b = (int)(tile / a.tiles_l), l0 = (int)(tile % a.tiles_l) * 128;
                const int s = it % STAGES, ph = (it / STAGES) & 1;
                bar_wait(&empty[s], ph ^ 1);
                bar_expect_tx(&full[s], a.rows * kRow);
                tma_load_3d(slabs + s * slab_stride, &map_a, &full[s], 0, l0 - a.halo, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(FMT_BF16, 128, C);
            bar_wait(wbar, 0);
            uint32_t it = 0;
            for (long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const uint32_t ab = it & 1, aph = (it >> 1) & 1;
                const int s = it % STAGES, ph = (it / STAGES) & 1;
                bar_wait(&tempty[ab], aph ^ 1);
                bar_wait(&full[s], ph);
                fence_after_sync();
                const uint32_t tmem_d = tmem_base + ab * C;
                const uint32_t slab = smem_addr(slabs + s * slab_stride);
                for (int tap = 0; tap < a.taps; ++tap) {
                    const uint64_t da = make_desc_k<C>(slab + tap * a.dil * kRow);
                    const uint64_t db = make_desc_k<C>(smem_addr(wsm + tap * kWTile));
#pragma unroll
                    for (int k = 0; k < C / 16; ++k) mma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (tap | k) != 0);
                }
                mma_commit(&empty[s]);
                mma_commit(&tfull[ab]);
            }
        }
    } else {
        const int quad = warp & 3;
        uint32_t it = 0;
        // Measured per layer at 256 clips (profiles/r02z*_vocoder_launches_b256.csv): the prefetching + staged epilogue wins
        // where the layer adds a residual and writes two outputs at 64 channels (1 123 -> 900 us, 47 -> 59 % of the DRAM
        // throughput) and loses everywhere else (more instructions in an epilogue that was not waiting on L1).
        if (a.epi_tma) {
            const EpiTma em{&map_or, &map_oa, &map_rs, a.res_tma};
            uint32_t rcount = 0;
            for (long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int b = (int)(tile / a.tiles_l), l0 = (int)(tile % a.tiles_l) * 128;
                const uint32_t ab = it & 1, aph = (it >> 1) & 1;
                if (a.res_tma)
                    conv_epilogue_tma<C, true>(a.e, em, b, l0, 0, tmem_base + ab * C, quad, lane, &tfull[ab], aph, stg, rbar, rcount);
                else
                    conv_epilogue_tma<C, false>(a.e, em, b, l0, 0, tmem_base + ab * C, quad, lane, &tfull[ab], aph, stg, rbar, rcount);
                fence_before_sync();
                __syncwarp();
                if (lane == 0) bar_arrive(&tempty[ab]);
            }
            if (lane == 0) bulk_wait0();
        } else if (C == 64 && a.e.resid != nullptr) {
            int4 res[C / 8];   // this thread's residual row of the tile about to be finished (requested a tile ahead)
            if ((long)blockIdx.x < total_tiles)
                epi_load_resid<C>(a.e, (int)(blockIdx.x / a.tiles_l), (int)(blockIdx.x % a.tiles_l) * 128, 0, quad, lane, true, res);
            for (long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int b = (int)(tile / a.tiles_l), l0 = (int)(tile % a.tiles_l) * 128;
                const uint32_t ab = it & 1, aph = (it >> 1) & 1;
                int4 cur[C / 8];
#pragma unroll
                for (int j = 0; j < C / 8; ++j) cur[j] = res[j];
                const long nt = tile + gridDim.x;
                epi_load_resid<C>(a.e, (int)(nt / a.tiles_l), (int)(nt % a.tiles_l) * 128, 0, quad, lane, nt < total_tiles, res);
                conv_epilogue_pre<C>(a.e, b, l0, 0, tmem_base + ab * C, quad, lane, &tfull[ab], aph, cur, stg);
                fence_before_sync();
                __syncwarp();
                if (lane == 0) bar_arrive(&tempty[ab]);
            }
        } else {
            for (long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int b = (int)(tile / a.tiles_l), l0 = (int)(tile % a.tiles_l) * 128;
                const uint32_t ab = it & 1, aph = (it >> 1) & 1;
                conv_epilogue<C>(a.e, b, l0, 0, tmem_base + ab * C, quad, lane, &tfull[ab], aph);
                fence_before_sync();
                __syncwarp();
                if (lane == 0) bar_arrive(&tempty[ab]);
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kCols);
}

// ---------------------------------------------------------------------------------------------------
// Wide-channel variant (C_in % 128 == 0, N % 128 == 0: upsample + MRF stages with 128 / 256 channels).
// With the general kernel these stages are L2 -> SM bandwidth bound: per 128-row tile it re-fetches a shifted
// A tile per tap AND the full weight matrix.  Here one work item is a PAIR of tiles (256 positions) x 128
// output channels: the activation slab (256 + 2*halo rows, two 64-channel blocks at a time) is fetched once
// and read through row-shifted descriptors, and every streamed 128x64 weight block feeds both tiles
// (two M = 128 MMAs into two TMEM accumulators), halving weight traffic per output.  Slab chunks and weight
// blocks ride separate mbarrier rings; accumulators are double-buffered (4 x 128 = 512 TMEM columns).
// ---------------------------------------------------------------------------------------------------
struct Slab2Args {
    EpiArgs e;
    int B, Cin, taps, dil, halo;
    int rb;                       // TMA box height: the slab (256 + 2*halo rows) arrives as two boxes
    int pairs_l, tiles_n, groups; // tile pairs per clip, 128-wide N tiles, channel groups of 128
    int epi_tma;                  // outputs through TMA stores (conv_epilogue_tma) instead of per-thread stores
    int res_tma;                  // ... and the residual block through TMA loads (8 KB of tiles per epilogue warp)
};

template <int WSTAGES>
__global__ void __launch_bounds__(kConvThreads, 1)
conv1d_slab2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                    const __grid_constant__ CUtensorMap map_or, const __grid_constant__ CUtensorMap map_oa,
                    const __grid_constant__ CUtensorMap map_rs, Slab2Args a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kRow = 128, kWTile = 128 * kRow;           // 64 channels bf16 per row; 128 x 64 weight block
    const int part = (2 * a.rb * kRow + 1023) & ~1023;       // one 64-channel block of the slab
    unsigned char* slabs = smem;                             // [2 buffers][2 channel blocks][part]
    unsigned char* wring = smem + 4 * part;                  // [WSTAGES][kWTile]
    uint64_t* sfull = reinterpret_cast<uint64_t*>(wring + WSTAGES * kWTile);
    uint64_t* sempty = sfull + 2;
    uint64_t* wfull = sempty + 2;
    uint64_t* wempty = wfull + WSTAGES;
    uint64_t* tfull = wempty + WSTAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    // per-warp tiles of the TMA epilogue (1 KB aligned: the 64-byte swizzle follows the address bits): raw, activated and -
    // with res_tma - two residual tiles; the residual tiles' mbarriers
    uint64_t* rbars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));
    unsigned char* stg = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(rbars + 8) + 1023) & ~uintptr_t(1023)) +
                         ((threadIdx.x >> 5) & 3) * (a.res_tma ? 8192 : 4096);
    uint64_t* rbar = rbars + 2 * ((threadIdx.x >> 5) & 3);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            bar_init(&sfull[i], 1);
            bar_init(&sempty[i], 1);
            bar_init(&tfull[i], 1);
            bar_init(&tempty[i], 4);
        }
        for (int s = 0; s < WSTAGES; ++s) {
            bar_init(&wfull[s], 1);
            bar_init(&wempty[s], 1);
        }
        for (int i = 0; i < 8; ++i) bar_init(&rbars[i], 1);
        bar_init_fence();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const long items = (long)a.B * a.pairs_l * a.tiles_n;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t ci = 0, wi = 0;
            for (long item = blockIdx.x; item < items; item += gridDim.x) {
                const int tn = (int)(item % a.tiles_n);
                const long pr = item / a.tiles_n;
                const int b = (int)(pr / a.pairs_l), l0 = (int)(pr % a.pairs_l) * 256;
                for (int g = 0; g < a.groups; ++g, ++ci) {
                    const int sb = ci & 1, sph = (ci >> 1) & 1;
                    bar_wait(&sempty[sb], sph ^ 1);
                    bar_expect_tx(&sfull[sb], 4 * a.rb * kRow);
                    for (int c = 0; c < 2; ++c)
                        for (int hf = 0; hf < 2; ++hf)
                            tma_load_3d(slabs + (2 * sb + c) * part + hf * a.rb * kRow, &map_a, &sfull[sb],
                                        (2 * g + c) * 64, l0 - a.halo + hf * a.rb, b);
                    for (int tap = 0; tap < a.taps; ++tap)
                        for (int c = 0; c < 2; ++c, ++wi) {
                            const int ws = wi % WSTAGES, wph = (wi / WSTAGES) & 1;
                            bar_wait(&wempty[ws], wph ^ 1);
                            bar_expect_tx(&wfull[ws], kWTile);
                            tma_load_2d(wring + ws * kWTile, &map_w, &wfull[ws], tap * a.Cin + (2 * g + c) * 64, tn * 128);
                        }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(FMT_BF16, 128, 128);
            uint32_t ci = 0, wi = 0, tcount = 0;
            for (long item = blockIdx.x; item < items; item += gridDim.x, ++tcount) {
                const uint32_t ab = tcount & 1, aph = (tcount >> 1) & 1;
                bar_wait(&tempty[ab], aph ^ 1);
                fence_after_sync();
                const uint32_t acc0 = tmem_base + ab * 256, acc1 = acc0 + 128;
                uint32_t first = 1;
                for (int g = 0; g < a.groups; ++g, ++ci) {
                    const int sb = ci & 1, sph = (ci >> 1) & 1;
                    bar_wait(&sfull[sb], sph);
                    fence_after_sync();
                    for (int tap = 0; tap < a.taps; ++tap)
                        for (int c = 0; c < 2; ++c, ++wi) {
                            const int ws = wi % WSTAGES, wph = (wi / WSTAGES) & 1;
                            bar_wait(&wfull[ws], wph);
                            fence_after_sync();
                            const uint32_t sa = smem_addr(slabs + (2 * sb + c) * part) + tap * a.dil * kRow;
                            const uint64_t da0 = make_desc_k<64>(sa), da1 = make_desc_k<64>(sa + 128 * kRow);
                            const uint64_t db = make_desc_k<64>(smem_addr(wring + ws * kWTile));
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                mma_f16(acc0, da0 + 2 * k, db + 2 * k, idesc, !(first && k == 0));
                                mma_f16(acc1, da1 + 2 * k, db + 2 * k, idesc, !(first && k == 0));
                            }
                            first = 0;
                            mma_commit(&wempty[ws]);
                        }
                    mma_commit(&sempty[sb]);
                }
                mma_commit(&tfull[ab]);
            }
        }
    } else {
        const int quad = warp & 3;
        uint32_t tcount = 0, rcount = 0;
        // (Measured and dropped for this kernel, profiles/r02z*_vocoder_launches_b256.csv: residual rows held in registers a
        // tile ahead - no change; the staged, coalesced write-out - 676 -> 820 us on the layers without a residual; eight
        // epilogue warps splitting the columns - 676 -> 931 us.  ncu shows L1 as the busiest unit (66 - 72 %) with 31.6 sectors
        // per store request, but every variant that traded those for more epilogue instructions lost.)
        for (long item = blockIdx.x; item < items; item += gridDim.x, ++tcount) {
            const int tn = (int)(item % a.tiles_n);
            const long pr = item / a.tiles_n;
            const int b = (int)(pr / a.pairs_l), l0 = (int)(pr % a.pairs_l) * 256;
            const uint32_t ab = tcount & 1, aph = (tcount >> 1) & 1;
            if (a.epi_tma) {
                const EpiTma em{&map_or, &map_oa, &map_rs, a.res_tma};
                if (a.res_tma) {
                    conv_epilogue_tma<128, true>(a.e, em, b, l0, tn, tmem_base + ab * 256, quad, lane, &tfull[ab], aph, stg, rbar, rcount);
                    conv_epilogue_tma<128, true>(a.e, em, b, l0 + 128, tn, tmem_base + ab * 256 + 128, quad, lane, &tfull[ab], aph, stg,
                                                 rbar, rcount);
                } else {
                    conv_epilogue_tma<128, false>(a.e, em, b, l0, tn, tmem_base + ab * 256, quad, lane, &tfull[ab], aph, stg, rbar, rcount);
                    conv_epilogue_tma<128, false>(a.e, em, b, l0 + 128, tn, tmem_base + ab * 256 + 128, quad, lane, &tfull[ab], aph, stg,
                                                  rbar, rcount);
                }
            } else {
                conv_epilogue<128>(a.e, b, l0, tn, tmem_base + ab * 256, quad, lane, &tfull[ab], aph);
                conv_epilogue<128>(a.e, b, l0 + 128, tn, tmem_base + ab * 256 + 128, quad, lane, &tfull[ab], aph);
            }
            fence_before_sync();
            __syncwarp();
            if (lane == 0) bar_arrive(&tempty[ab]);
        }
        if (a.epi_tma && lane == 0) bulk_wait0();   // the tiles must outlive the last stores' reads; the data is then in flight to L2
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int g_conv_epi_tma = 1;   // adv_set_conv_epilogue(): 1 = TMA-store epilogue where it exists (default), 0 = per-thread stores

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

static int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <class K>
static int set_smem_attr2(K kernel, size_t bytes) {
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

// resident CTAs per SM for a persistent launch (registers, shared memory and TMEM columns all bound it)
template <class K>
static int resident_ctas(K kernel, size_t smem, int tmem_cols) {
    static std::mutex mu;
    static std::unordered_map<size_t, int> cache;
    const size_t key = reinterpret_cast<size_t>(kernel) ^ (smem * 1315423911u);
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    const int n = adv_resident_ctas(kernel, kConvThreads, smem, tmem_cols, 4);
    cache[key] = n;
    return n;
}

template <int BN, int BK, int STAGES>
static int launch_conv_tma(const CUtensorMap& ma, const CUtensorMap& mw, ConvTmaArgs& a, cudaStream_t s) {
    constexpr size_t smem = (size_t)STAGES * (128 * BK * 2 + BN * BK * 2) + 256 + 1024;
    int rc = set_smem_attr2(conv1d_tma_kernel<BN, BK, STAGES>, smem);
    if (rc != ADV_OK) return rc;
    a.tiles_n = a.N / BN;
    const long tiles = (long)a.B * a.tiles_l * a.tiles_n;
    const int per_sm = resident_ctas(conv1d_tma_kernel<BN, BK, STAGES>, smem, 2 * BN);
    long grid = (long)num_sms() * per_sm;
    if (grid > tiles) grid = tiles;
    conv1d_tma_kernel<BN, BK, STAGES><<<(unsigned)grid, kConvThreads, smem, s>>>(ma, mw, a);
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

template <int C>
static int launch_conv_slab(const CUtensorMap& ma, const CUtensorMap& mw, const CUtensorMap& m_or, const CUtensorMap& m_oa,
                            const CUtensorMap& m_rs, const SlabArgs& a, cudaStream_t s) {
    constexpr int STAGES = 3;
    const size_t slab_stride = ((size_t)a.rows * C * 2 + 1023) & ~size_t(1023);
    const size_t smem = (size_t)a.taps * C * C * 2 + STAGES * slab_stride + 256 + 1024 + 4 * 8192 + 1024 + 128;
    int rc = set_smem_attr2(conv1d_slab_kernel<C, STAGES>, smem);
    if (rc != ADV_OK) return rc;
    const int per_sm = resident_ctas(conv1d_slab_kernel<C, STAGES>, smem, 2 * C);
    const long tiles = (long)a.B * a.tiles_l;
    long grid = (long)num_sms() * per_sm;
    if (grid > tiles) grid = tiles;
    conv1d_slab_kernel<C, STAGES><<<(unsigned)grid, kConvThreads, smem, s>>>(ma, mw, m_or, m_oa, m_rs, a);
    ADV_CUDA_CHECK(cudaGetLastError());
    return ADV_OK;
}

}  // namespace adv

using namespace adv;

extern "C" int adv_set_conv_epilogue(int mode) {   // 0: per-thread stores; 1: TMA stores + TMA residual loads; 2: TMA stores only
    const int prev = adv::g_conv_epi_tma;
    adv::g_conv_epi_tma = mode < 0 || mode > 2 ? 1 : mode;
    return prev;
}

extern "C" int adv_conv1d_bf16_tma(const void* in, const void* w, const float* bias, const void* resid, void* out_raw,
                                   void* out_act, int batch, int L, int Cin, int taps, int dil, int N, float act_slope,
                                   float out_scale, void* stream) {
    if (!in || !w || (!out_raw && !out_act) || batch <= 0 || L <= 0 || taps <= 0 || dil <= 0 || N <= 0) return ADV_ERR_INVALID;
    if ((taps & 1) == 0 || N % 32 != 0 || !(Cin == 32 || Cin % 64 == 0)) return ADV_ERR_SHAPE;
    EncodeTiledFn enc = encode_fn();
    if (!enc) return ADV_ERR_UNSUPPORTED;
    const int BK = Cin == 32 ? 32 : 64;
    const int BN = N % 256 == 0 ? 256 : (N % 128 == 0 ? 128 : (N % 64 == 0 ? 64 : 32));
    const CUtensorMapSwizzle swz = BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;

    CUtensorMap ma, mw;
    {   // activations [B][L][Cin]: dims innermost first
        cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)L, (cuuint64_t)batch};
        cuuint64_t strides[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)L * Cin * 2};
        cuuint32_t box[3] = {(cuuint32_t)BK, 128, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(in), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return ADV_ERR_INVALID;
    }
    {   // weights [N][taps*Cin]
        cuuint64_t dims[2] = {(cuuint64_t)taps * Cin, (cuuint64_t)N};
        cuuint64_t strides[1] = {(cuuint64_t)taps * Cin * 2};
        cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
        cuuint32_t estr[2] = {1, 1};
        if (enc(&mw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return ADV_ERR_INVALID;
    }
    static const bool no_slab = ADV_AB_ENV("ADV_NO_SLAB") != nullptr;
    if (!no_slab && Cin == N && (Cin == 32 || Cin == 64)) {
        const int halo = ((taps - 1) / 2) * dil, rows = 128 + 2 * halo;
        if (rows <= 256) {
            CUtensorMap ms;  // slab map: box = all C channels x (128 + 2 halo) rows x 1 clip
            cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)L, (cuuint64_t)batch};
            cuuint64_t strides[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)L * Cin * 2};
            cuuint32_t box[3] = {(cuuint32_t)Cin, (cuuint32_t)rows, 1};
            cuuint32_t estr[3] = {1, 1, 1};
            if (enc(&ms, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(in), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return ADV_ERR_INVALID;
            CUtensorMap mws;  // one tap of the weights: box = C (k) x C (n)
            cuuint64_t wd[2] = {(cuuint64_t)taps * Cin, (cuuint64_t)N};
            cuuint64_t wst[1] = {(cuuint64_t)taps * Cin * 2};
            cuuint32_t wb[2] = {(cuuint32_t)Cin, (cuuint32_t)N};
            cuuint32_t we[2] = {1, 1};
            if (enc(&mws, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), wd, wst, wb, we,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return ADV_ERR_INVALID;
            SlabArgs sa;
            sa.e = EpiArgs{bias, (const __nv_bfloat16*)resid, (__nv_bfloat16*)out_raw, (__nv_bfloat16*)out_act, L, N,
                           act_slope, out_scale};
            sa.B = batch; sa.taps = taps; sa.dil = dil; sa.halo = halo; sa.rows = rows;
            sa.tiles_l = (L + 127) / 128;
            CUtensorMap m_or = ms, m_oa = ms, m_rs = ms;   // output / residual maps of the TMA epilogue: [N][L][B], box 32 x 32, 64-byte swizzle
            // tensor-bound layers keep the per-thread stores: an 11-tap conv without a residual and with one output loses 4 - 6 %
            // to the staged form (its four epilogue warps have no slack), measured per layer in profiles/r02zh_* / r02zk_*
            const bool mma_heavy = taps >= 11 && resid == nullptr && !(out_raw && out_act);
            sa.epi_tma = g_conv_epi_tma != 0 && !mma_heavy;
            sa.res_tma = sa.epi_tma && resid != nullptr && g_conv_epi_tma != 2;
            for (int k = 0; k < 3 && sa.epi_tma; ++k) {
                void* dst = k == 0 ? out_raw : (k == 1 ? out_act : const_cast<void*>(resid));
                if (!dst || (k == 2 && !sa.res_tma)) continue;
                cuuint64_t od[3] = {(cuuint64_t)N, (cuuint64_t)L, (cuuint64_t)batch};
                cuuint64_t ost[2] = {(cuuint64_t)N * 2, (cuuint64_t)L * N * 2};
                cuuint32_t ob[3] = {32, 32, 1};
                cuuint32_t oe[3] = {1, 1, 1};
                if (enc(k == 0 ? &m_or : (k == 1 ? &m_oa : &m_rs), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dst, od, ost, ob, oe,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                    sa.epi_tma = sa.res_tma = 0;
            }
            return Cin == 64 ? launch_conv_slab<64>(ms, mws, m_or, m_oa, m_rs, sa, (cudaStream_t)stream)
                             : launch_conv_slab<32>(ms, mws, m_or, m_oa, m_rs, sa, (cudaStream_t)stream);
        }
    }
    static const bool no_slab2 = ADV_AB_ENV("ADV_NO_SLAB2") != nullptr;
    if (!no_slab && !no_slab2 && Cin % 128 == 0 && N % 128 == 0) {
        const int halo = ((taps - 1) / 2) * dil, rows = 256 + 2 * halo, rb = (rows + 1) / 2;
        if (rb <= 256) {
            CUtensorMap ms, mws;
            cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)L, (cuuint64_t)batch};
            cuuint64_t strides[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)L * Cin * 2};
            cuuint32_t box[3] = {64, (cuuint32_t)rb, 1};
            cuuint32_t estr[3] = {1, 1, 1};
            if (enc(&ms, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(in), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return ADV_ERR_INVALID;
            cuuint64_t wd[2] = {(cuuint64_t)taps * Cin, (cuuint64_t)N};
            cuuint64_t wst[1] = {(cuuint64_t)taps * Cin * 2};
            cuuint32_t wb[2] = {64, 128};
            cuuint32_t we[2] = {1, 1};
            if (enc(&mws, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), wd, wst, wb, we,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return ADV_ERR_INVALID;
            Slab2Args sa;
            sa.e = EpiArgs{bias, (const __nv_bfloat16*)resid, (__nv_bfloat16*)out_raw, (__nv_bfloat16*)out_act, L, N,
                           act_slope, out_scale};
            sa.B = batch; sa.Cin = Cin; sa.taps = taps; sa.dil = dil; sa.halo = halo; sa.rb = rb;
            sa.pairs_l = (L + 255) / 256; sa.tiles_n = N / 128; sa.groups = Cin / 128;
            constexpr int WS = 3;
            const size_t part = ((size_t)2 * rb * 128 + 1023) & ~size_t(1023);
            size_t smem = 4 * part + (size_t)WS * 128 * 128 + 256 + 1024;
            // output maps of the TMA-store epilogue: [N][L][B], box 32 channels x 32 rows, 64-byte swizzle
            CUtensorMap m_or = ms, m_oa = ms, m_rs = ms;
            // (the 256-channel MRF convs - not the transposed convs - and the 11-tap convs without a residual are tensor-bound)
            const bool mma_heavy = (Cin == 256 && N == 256) || (taps >= 11 && resid == nullptr && !(out_raw && out_act));
            sa.epi_tma = g_conv_epi_tma != 0 && !mma_heavy && smem + 4 * 4096 + 1024 + 128 <= 227 * 1024;
            sa.res_tma = sa.epi_tma && resid != nullptr && g_conv_epi_tma != 2 && smem + 4 * 8192 + 1024 + 128 <= 227 * 1024;
            for (int k = 0; k < 3 && sa.epi_tma; ++k) {
                void* dst = k == 0 ? out_raw : (k == 1 ? out_act : const_cast<void*>(resid));
                if (!dst || (k == 2 && !sa.res_tma)) continue;
                cuuint64_t od[3] = {(cuuint64_t)N, (cuuint64_t)L, (cuuint64_t)batch};
                cuuint64_t ost[2] = {(cuuint64_t)N * 2, (cuuint64_t)L * N * 2};
                cuuint32_t ob[3] = {32, 32, 1};
                cuuint32_t oe[3] = {1, 1, 1};
                if (enc(k == 0 ? &m_or : (k == 1 ? &m_oa : &m_rs), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dst, od, ost, ob, oe,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                    sa.epi_tma = sa.res_tma = 0;
            }
            if (sa.epi_tma) smem += 4 * (sa.res_tma ? 8192 : 4096) + 1024 + 128;
            int rc = set_smem_attr2(conv1d_slab2_kernel<WS>, smem);
            if (rc == ADV_OK) {
                const long items = (long)batch * sa.pairs_l * sa.tiles_n;
                long grid = num_sms();
                if (grid > items) grid = items;
                conv1d_slab2_kernel<WS><<<(unsigned)grid, kConvThreads, smem, (cudaStream_t)stream>>>(ms, mws, m_or, m_oa, m_rs, sa);
                ADV_CUDA_CHECK(cudaGetLastError());
                return ADV_OK;
            }
        }
    }
    ConvTmaArgs a;
    a.bias = bias;
    a.resid = (const __nv_bfloat16*)resid;
    a.out_raw = (__nv_bfloat16*)out_raw;
    a.out_act = (__nv_bfloat16*)out_act;
    a.B = batch; a.L = L; a.Cin = Cin; a.taps = taps; a.dil = dil; a.center = (taps - 1) / 2; a.N = N;
    a.tiles_l = (L + 127) / 128; a.tiles_n = 0;
    a.act_slope = act_slope; a.out_scale = out_scale;
    cudaStream_t s = (cudaStream_t)stream;
    if (BK == 64) {
        switch (BN) {
            case 256: return launch_conv_tma<256, 64, 4>(ma, mw, a, s);
            case 128: return launch_conv_tma<128, 64, 3>(ma, mw, a, s);
            case 64: return launch_conv_tma<64, 64, 4>(ma, mw, a, s);
            default: return launch_conv_tma<32, 64, 4>(ma, mw, a, s);
        }
    }
    switch (BN) {
        case 256: return launch_conv_tma<256, 32, 4>(ma, mw, a, s);
        case 128: return launch_conv_tma<128, 32, 4>(ma, mw, a, s);
        case 64: return launch_conv_tma<64, 32, 4>(ma, mw, a, s);
        default: return launch_conv_tma<32, 32, 4>(ma, mw, a, s);
    }
}
