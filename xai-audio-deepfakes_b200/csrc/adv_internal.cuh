// Internal definitions shared by the kernel translation units of libaddvisor_sm100.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/addvisor_b200.h"

#include <stdlib.h>
// A/B switches (environment variables selecting an older kernel generation or an alternative policy) exist only in
// -DADV_AB builds (ADV_NVCC_EXTRA=-DADV_AB); the product build reads no environment variable on a launch path and does
// not instantiate the superseded kernels.
#ifdef ADV_AB
#define ADV_AB_ENV(name) getenv(name)
#else
#define ADV_AB_ENV(name) ((const char*)nullptr)
#endif

namespace adv {

constexpr int kThreads = 256;  // CTA size of the transform kernels

// Plan fields the kernels read (passed by value as a kernel argument).
struct PlanDev {
    int n_fft, hop, T, n_in, n_out;
    int wlo, whi;          // window support [wlo, whi) inside n_fft (non-zero taps)
    int phases;            // ceil(support / hop): frames t and t+phases never overlap
    int rect_full;         // window == 1 on all n_fft taps (torch's default when win_length == n_fft)
    const float* window;   // dev [n_fft], window centred in n_fft
    const float* inv_env;  // dev [n_out], 1 / (n_fft * sum_t w^2), 0 where no frame lands
    const float2* tw;      // dev [2][32][lanes]: exp(-2*pi*i*l*k1/n_fft), then the row-rotated variant
    const float2* tw3;     // dev [f3::TW_TOTAL]: twiddle tables of the generation-3 warp FFT (fft3.cuh)
};

// Tiling of the output (sample) axis for the overlap-add kernels.
struct Tiling {
    int hops_per_tile;  // output samples per tile = hops_per_tile * hop
    int tiles;          // tiles per clip
};

}  // namespace adv

struct adv_plan {
    adv::PlanDev d;
    int win_length;
    int frames_per_tile;  // capacity of one CTA pass: 2 * units
    int max_hops;         // largest hops_per_tile whose frame span fits frames_per_tile
    int max_hops_cap[2];  // the same for tiles of 16 / 32 frames (generation-3 kernels: 16 warps x 1 or 2 frames)
    int rect_sup;         // window == 1 on its whole support [wlo, whi) (a rectangular win_length < n_fft window)
    int gen3;             // generation-3 kernels usable for this geometry (n_fft 512, or n_fft 1024 with an even hop)
    int device;
    void* dev_block;      // single allocation holding window / inv_env / tw / work counters
    int* work_ctr;        // dev [kWorkSlots][kWorkInts] zeroed ints: draw counters + finished warps of a dynamically scheduled launch
    unsigned work_next;   // host: slot of the next such launch (rotates, so launches in flight never share a slot)
};

namespace adv {
constexpr int kWorkSlots = 64, kWorkGroups = 16, kWorkPad = 64, kWorkInts = (kWorkGroups + 1) * kWorkPad;   // per launch: kWorkGroups draw counters + the finished-warp count, 256 bytes apart (separate L2 lines)
// Counter pair for one launch of a dynamically scheduled kernel (stft3_kernel).  The kernel leaves both at zero when its
// last warp retires, so a slot can be baked into a CUDA graph and replayed; slots rotate per launch, so up to kWorkSlots
// launches of one plan may be in flight (or captured in graphs that run concurrently) at once.
static inline int* next_work_slot(const adv_plan* p) {
    const unsigned k = __atomic_fetch_add(&const_cast<adv_plan*>(p)->work_next, 1u, __ATOMIC_RELAXED);
    return p->work_ctr + kWorkInts * (k % kWorkSlots);
}
}  // namespace adv

namespace adv {
void set_cuda_error(cudaError_t e);
// slots_per_sm: resident CTAs per SM of the consumer; frames_cap: frames one CTA pass holds (0: the plan's default)
Tiling choose_tiling(const adv_plan* p, int batch, int slots_per_sm, bool balanced = false, int frames_cap = 0);
// generation-3 launchers (transform3_kernels.cu); return ADV_ERR_UNSUPPORTED when the call is outside their domain
bool gen3_enabled();
int launch_stft3(const adv_plan* p, const float* wav, int64_t wav_stride, int batch, float2* X, float* mag,
                 float* phase, int flags, cudaStream_t s);
int launch_istft3(const adv_plan* p, const float2* X, int64_t sb, int64_t st, int64_t sf, int batch, float* out,
                  double* stats, cudaStream_t s);
int launch_explain3(const adv_plan* p, const float* wav, int64_t wav_stride, const float2* X, int64_t sb, int64_t st,
                    int64_t sf, const float* mask, int Fm, int Tm, int mode, int batch, float* rel, float* irr,
                    double* stats, cudaStream_t s);
// generation-4 streaming explain kernel (transform4_kernels.cu): n_fft 512, rectangular full window, hop 128 / 160 / 256,
// waveform input.  explain4_slots(): statistics slots per clip of such a launch, 0 when the plan is outside its domain.
int launch_explain4(const adv_plan* p, const float* wav, int64_t wav_stride, const float* mask, int Fm, int Tm,
                    int mode_flags, int batch, float* rel, float* irr, double* stats, cudaStream_t s);
int explain4_slots(const adv_plan* p, int batch);
// streaming iSTFT of the same generation (same domain, any spectrum strides); istft4_slots(): statistics slots per clip
int launch_istft4(const adv_plan* p, const float2* X, int64_t sb, int64_t st, int64_t sf, int batch, float* out,
                  double* stats, cudaStream_t s);
int istft4_slots(const adv_plan* p, int batch);
// streaming kernels for n_fft 1024 (transform5_kernels.cu): even hop / window support / n_out, 64 <= hop <= 512, support <= 4 hops
int launch_istft5(const adv_plan* p, const float2* X, int64_t sb, int64_t st, int64_t sf, int batch, float* out,
                  double* stats, cudaStream_t s);
int istft5_slots(const adv_plan* p, int batch);
int launch_stft5(const adv_plan* p, const float* wav, int64_t wav_stride, int batch, float2* X, float* mag, float* phase,
                 int flags, cudaStream_t s);
int launch_explain5(const adv_plan* p, const float* wav, int64_t wav_stride, const float* mask, int Fm, int Tm,
                    int mode_flags, int batch, float* rel, float* irr, double* stats, cudaStream_t s);
int explain5_slots(const adv_plan* p, int batch);   // 0: outside the kernel's domain (or its strips / slices do not fit)
int istft3_frames_cap(const adv_plan* p);    // 32 when launch_istft3 takes the plan, else 0 (plan default)
bool istft_balanced();  // tiling policy of the stand-alone iSTFT kernels (ADV_ISTFT_BALANCED=1 selects the round-balanced tile length; default: longest tile)
int launch_stft(const adv_plan* p, const float* wav, int64_t wav_stride, int batch, float2* X, float* mag,
                float* phase, int flags, cudaStream_t s);
int launch_istft(const adv_plan* p, const float2* X, int64_t sb, int64_t st, int64_t sf, int batch, float* out,
                 double* stats, cudaStream_t s);
int launch_explain(const adv_plan* p, const float* wav, int64_t wav_stride, const float2* X, int64_t sb,
                   int64_t st, int64_t sf, const float* mask, int Fm, int Tm, int mode, int batch, float* rel,
                   float* irr, double* stats, cudaStream_t s);
}  // namespace adv

// Co-resident CTAs per SM of a persistent kernel, from the resources themselves (registers as compiled, dynamic +
// static + 1 KB reserved shared memory per CTA out of 228 KB, TMEM columns out of 512).  The occupancy API was
// measured to answer 1 for these large-dynamic-shared-memory kernels on this driver, which halved the grids.
template <class K>
static inline int adv_resident_ctas(K kernel, int threads, size_t dyn_smem, int tmem_cols, int cap) {
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) return 1;
    const int warps = (threads + 31) / 32;
    const int regs_per_warp = ((fa.numRegs + 7) / 8) * 8 * 32;
    int n = regs_per_warp > 0 ? 65536 / (regs_per_warp * warps) : cap;
    const size_t per_cta = dyn_smem + fa.sharedSizeBytes + 1024;
    const int by_smem = (int)((size_t)228 * 1024 / per_cta);
    if (n > by_smem) n = by_smem;
    if (tmem_cols > 0) {
        int cols = 32;
        while (cols < tmem_cols) cols <<= 1;
        if (n > 512 / cols) n = 512 / cols;
    }
    const int by_threads = 2048 / (warps * 32);
    if (n > by_threads) n = by_threads;
    if (n > cap) n = cap;
    return n < 1 ? 1 : n;
}

// Programmatic dependent launch (sm_90+): a kernel launched through launch_pdl() may start its CTAs - table
// staging, barrier initialisation, index arithmetic - while the previous kernel of the stream is still draining
// its last round; pdl_wait() (griddepcontrol.wait) then blocks until that kernel has completed and its memory is
// visible, and MUST precede the first access to anything another kernel may have produced or may still read
// (inputs and outputs alike; plan tables are constants).  pdl_launch_dependents() lets the next kernel of the
// stream do the same with us.  Works in stream capture (programmatic graph edges).  Opt-in (ADV_PDL=1): see
// pdl_enabled() in capi.cu for the measurement behind the default.
namespace adv {
bool pdl_enabled();
template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                     Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
}  // namespace adv

#define ADV_CUDA_CHECK(expr)                         \
    do {                                             \
        cudaError_t _e = (expr);                     \
        if (_e != cudaSuccess) {                     \
            adv::set_cuda_error(_e);                 \
            return ADV_ERR_CUDA;                     \
        }                                            \
    } while (0)
