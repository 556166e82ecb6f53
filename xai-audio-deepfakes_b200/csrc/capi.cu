// C-ABI entry points: plan management, argument validation, launch dispatch.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "adv_internal.cuh"
#include "fft_core.cuh"
#include "fft3.cuh"

namespace adv {

static thread_local char g_cuda_err[256] = "";
void set_cuda_error(cudaError_t e) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
    (void)cudaGetLastError();
}

static int floordiv_h(int a, int b) {
    int q = a / b;
    return (a % b != 0 && a < 0) ? q - 1 : q;
}
static int ceildiv_h(int a, int b) { return floordiv_h(a + b - 1, b); }

// frames a tile of `k` hops needs, maximum over all tiles (same formulas as tile_geom() on the device)
static int tile_frame_span(const PlanDev& d, int k) {
    const int S = k * d.hop, half = d.n_fft / 2;
    int worst = 0;
    for (int s0 = 0; s0 < d.n_out; s0 += S) {
        const int s1 = s0 + S < d.n_out ? s0 + S : d.n_out;
        const int p0 = s0 + half, p1 = s1 + half;
        int t_lo = ceildiv_h(p0 - d.whi + 1, d.hop);
        int t_hi = floordiv_h(p1 - 1 - d.wlo, d.hop);
        if (t_lo < 0) t_lo = 0;
        if (t_hi > d.T - 1) t_hi = d.T - 1;
        const int span = t_hi - t_lo + 1;
        if (span > worst) worst = span;
    }
    return worst;
}

// Programmatic dependent launch: on by default (adv_set_pdl).  Measured (B200, 64 x 4 s clips): back-to-back launches on
// ONE stream gain 0.4 - 1.2 us each (the next kernel's table staging and barrier set-up run under the previous kernel's
// tail), but the pooled multi-stream step of pipeline.PipelinedPool LOSES 20 % (774 k -> 613 k clips/s): early-scheduled
// CTAs of the next explain kernel sit in griddepcontrol.wait on the shared memory the normaliser CTAs of the
// high-priority stream were meant to use - that pipeline switches it off for its own capture.
static int g_pdl = 1;
bool pdl_enabled() { return __atomic_load_n(&g_pdl, __ATOMIC_RELAXED) != 0; }

bool istft_balanced() {
    static const char* e = ADV_AB_ENV("ADV_ISTFT_BALANCED");
    static const bool on = e && e[0] == '1';  // measured: no gain for the wide-unit kernel (30.7 vs 29.7 us), off
    return on;
}

bool gen3_enabled() {
    static const char* e = ADV_AB_ENV("ADV_GEN3");   // A/B switch: ADV_GEN3=0 routes every call to the generation-2 kernels
    static const bool on = !(e && e[0] == '0');
    return on;
}

Tiling choose_tiling(const adv_plan* p, int batch, int slots_per_sm, bool balanced, int frames_cap) {
    // A tile costs one CTA pass whose length grows with the frames it transforms (each unit with a live frame
    // runs its FFTs; idle units skip them), and the persistent kernels run ceil(tiles * batch / resident CTAs)
    // rounds.  Pick the tile length that minimises rounds x (frames + fixed cost) instead of simply the longest
    // tile: for 64 clips of 400 hops the longest tile (29 hops, 32 frames) gives 896 tiles = 6.05 rounds on 148
    // one-CTA SMs but costs 7 full rounds, while 25 hops (28 frames) gives 1024 tiles = 6.92 rounds of shorter
    // passes - the same number of frames transformed, 12 % less time.
    int n_sm = 0;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        static int cache[64] = {0};
        if (dev >= 0 && dev < 64 && cache[dev] > 0) n_sm = cache[dev];
        else {
            if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
            if (dev >= 0 && dev < 64) cache[dev] = n_sm;
        }
    }
    const PlanDev& d = p->d;
    const int total_hops = (d.n_out + d.hop - 1) / d.hop;
    const long slots = (long)n_sm * (slots_per_sm < 1 ? 1 : slots_per_sm);
    // Measured (64 x 4 s clips): the balanced length makes the fused kernel ALONE 2.7 % faster (88.9 vs 91.4 us) but the
    // pipelined step 9 % slower (101.7 vs 94.8 us): with the longest tiles the 7th round is almost empty and the next
    // kernels / the next batch's launch fill it, while 6.92 full rounds leave nothing to overlap.  The longest tile is
    // therefore the default; ADV_TILING_BALANCED=1 selects the balanced policy (stand-alone kernel latency).
    static const bool env_longest = ADV_AB_ENV("ADV_TILING_BALANCED") == nullptr;
    // (the stand-alone iSTFT can ask for the balanced policy, see istft_balanced)
    const bool longest = env_longest && !balanced;
    int best_k = 0;
    long best_cost = 0;
    const int max_hops = frames_cap == 16 ? p->max_hops_cap[0] : (frames_cap == 32 ? p->max_hops_cap[1] : p->max_hops);
    const int k_min = longest ? max_hops : (max_hops / 2 > 1 ? max_hops / 2 : 1);
    for (int k0 = max_hops; k0 >= k_min; --k0) {
        const int tiles = (total_hops + k0 - 1) / k0;
        const int k = (total_hops + tiles - 1) / tiles;  // even split of the same number of tiles
        const long rounds = ((long)tiles * (batch < 1 ? 1 : batch) + slots - 1) / slots;
        const int frames = (tile_frame_span(d, k) + 1) & ~1;
        const long cost = rounds * (frames + 8);
        if (best_k == 0 || cost < best_cost) {
            best_k = k;
            best_cost = cost;
        }
    }
    return Tiling{best_k, (total_hops + best_k - 1) / best_k};
}

}  // namespace adv

using namespace adv;

extern "C" {

int adv_version(void) { return 101; }

int adv_set_pdl(int on) { return __atomic_exchange_n(&adv::g_pdl, on ? 1 : 0, __ATOMIC_RELAXED); }

const char* adv_strerror(int status) {
    switch (status) {
        case ADV_OK: return "ok";
        case ADV_ERR_INVALID: return "invalid argument";
        case ADV_ERR_UNSUPPORTED: return "unsupported geometry (n_fft must be 512 or 1024, hop <= n_fft)";
        case ADV_ERR_NOLA: return "window overlap add min: 1";
        case ADV_ERR_SHORT_INPUT: return "input shorter than the reflect padding (n_in must exceed n_fft/2)";
        case ADV_ERR_CUDA: return "CUDA error";
        case ADV_ERR_SHAPE: return "shape mismatch with the plan";
        default: return "unknown status";
    }
}

const char* adv_last_cuda_error(void) { return g_cuda_err; }

int adv_plan_create(adv_plan** out, int n_fft, int hop, int win_length, const float* window_host, int n_frames,
                    int n_in, int n_out) {
    if (!out) return ADV_ERR_INVALID;
    *out = nullptr;
    if (n_fft <= 0 || hop <= 0 || win_length <= 0 || win_length > n_fft || n_frames <= 0 || n_out < 0)
        return ADV_ERR_INVALID;
    if (n_fft != 512 && n_fft != 1024) return ADV_ERR_UNSUPPORTED;
    if (n_in > 0 && n_in <= n_fft / 2) return ADV_ERR_SHORT_INPUT;
    if (n_in > 0 && 1 + n_in / hop != n_frames) return ADV_ERR_SHAPE;

    // window centred in n_fft (torch.stft pads win_length < n_fft on both sides)
    std::vector<float> win(n_fft, 0.0f);
    const int left = (n_fft - win_length) / 2;
    for (int i = 0; i < win_length; ++i) win[left + i] = window_host ? window_host[i] : 1.0f;
    int wlo = 0, whi = n_fft;
    while (wlo < n_fft && win[wlo] == 0.0f) ++wlo;
    while (whi > wlo && win[whi - 1] == 0.0f) --whi;
    if (whi <= wlo) return ADV_ERR_NOLA;
    const int support = whi - wlo;
    const int phases = (support + hop - 1) / hop;
    if (phases > 64) return ADV_ERR_UNSUPPORTED;

    // overlap-add envelope over the kept span [n_fft/2, n_fft/2 + n_out) and torch.istft's NOLA check
    const long full_len = (long)n_fft + (long)hop * (n_frames - 1);
    std::vector<double> env(full_len, 0.0);
    for (int t = 0; t < n_frames; ++t)
        for (int n = wlo; n < whi; ++n) env[(long)t * hop + n] += (double)win[n] * (double)win[n];
    std::vector<float> inv_env(n_out, 0.0f);
    for (int s = 0; s < n_out; ++s) {
        const long p = (long)s + n_fft / 2;
        if (p >= full_len) break;  // torch zero-pads the tail when `length` exceeds the signal
        if (fabs(env[p]) < 1e-11) return ADV_ERR_NOLA;
        inv_env[s] = (float)(1.0 / ((double)n_fft * env[p]));
    }

    const int lanes = n_fft / 32;
    // [0]: exp(-2 pi i l k1 / n_fft); [1]: the same times exp(-2 pi i k1 / 32), the table of a unit that reads
    // its samples rotated by one radix-32 row (stft_w_kernel's bank-conflict avoidance)
    std::vector<float2> tw((size_t)2 * 32 * lanes);
    for (int r = 0; r < 2; ++r)
        for (int k1 = 0; k1 < 32; ++k1)
            for (int l = 0; l < lanes; ++l) {
                const long num = ((long)k1 * l + (long)r * k1 * (n_fft / 32)) % n_fft;
                const double a = -2.0 * M_PI * (double)num / (double)n_fft;
                tw[((size_t)r * 32 + k1) * lanes + l] = make_float2((float)cos(a), (float)sin(a));
            }

    adv_plan* p = (adv_plan*)calloc(1, sizeof(adv_plan));
    if (!p) return ADV_ERR_INVALID;
    std::vector<float2> tw3(f3::TW_TOTAL);
    f3::build_tables(tw3.data());
    const size_t b_win = sizeof(float) * n_fft, b_env = sizeof(float) * (size_t)n_out, b_tw = sizeof(float2) * tw.size();
    const size_t b_tw3 = sizeof(float2) * tw3.size();
    const size_t o_env = (b_win + 255) / 256 * 256, o_tw = o_env + (b_env + 255) / 256 * 256;
    const size_t o_tw3 = o_tw + (b_tw + 255) / 256 * 256;
    const size_t o_work = o_tw3 + (b_tw3 + 255) / 256 * 256, b_work = sizeof(int) * kWorkInts * kWorkSlots;
    cudaError_t e = cudaMalloc(&p->dev_block, o_work + b_work);
    if (e == cudaSuccess) e = cudaMemset((char*)p->dev_block + o_work, 0, b_work);
    if (e == cudaSuccess) e = cudaMemcpy((char*)p->dev_block + o_tw3, tw3.data(), b_tw3, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy((char*)p->dev_block, win.data(), b_win, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy((char*)p->dev_block + o_env, inv_env.data(), b_env, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy((char*)p->dev_block + o_tw, tw.data(), b_tw, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        set_cuda_error(e);
        if (p->dev_block) cudaFree(p->dev_block);
        free(p);
        return ADV_ERR_CUDA;
    }
    cudaGetDevice(&p->device);
    p->d.n_fft = n_fft;
    p->d.hop = hop;
    p->d.T = n_frames;
    p->d.n_in = n_in;
    p->d.n_out = n_out;
    p->d.wlo = wlo;
    p->d.whi = whi;
    p->d.phases = phases;
    p->d.rect_full = 1;
    for (int i = 0; i < n_fft; ++i)
        if (win[i] != 1.0f) p->d.rect_full = 0;
    p->rect_sup = 1;
    for (int i = wlo; i < whi; ++i)
        if (win[i] != 1.0f) p->rect_sup = 0;
    p->d.window = (const float*)p->dev_block;
    p->d.inv_env = (const float*)((char*)p->dev_block + o_env);
    p->d.tw = (const float2*)((char*)p->dev_block + o_tw);
    p->d.tw3 = (const float2*)((char*)p->dev_block + o_tw3);
    p->work_ctr = (int*)((char*)p->dev_block + o_work);
    p->work_next = 0;
    p->gen3 = (n_fft == 512 || (n_fft == 1024 && hop % 2 == 0)) ? 1 : 0;
    p->win_length = win_length;
    p->frames_per_tile = 2 * (kThreads / lanes);
    p->max_hops = n_out == 0 ? 1 : 0;  // n_out == 0: forward-only plan, no overlap-add tiling
    for (int k = p->frames_per_tile; k >= 1 && n_out > 0; --k)
        if (tile_frame_span(p->d, k) <= p->frames_per_tile) {
            p->max_hops = k;
            break;
        }
    for (int c = 0; c < 2; ++c) {
        const int cap = c == 0 ? 16 : 32;
        p->max_hops_cap[c] = 0;
        for (int k = cap; k >= 1 && n_out > 0; --k)
            if (tile_frame_span(p->d, k) <= cap) {
                p->max_hops_cap[c] = k;
                break;
            }
    }
    if (p->max_hops == 0) {  // hop so small that even a one-hop tile needs more frames than a CTA holds
        adv_plan_destroy(p);
        return ADV_ERR_UNSUPPORTED;
    }
    *out = p;
    return ADV_OK;
}

void adv_plan_destroy(adv_plan* plan) {
    if (!plan) return;
    if (plan->dev_block) cudaFree(plan->dev_block);
    free(plan);
}

int adv_plan_bins(const adv_plan* plan) { return plan ? plan->d.n_fft / 2 + 1 : ADV_ERR_INVALID; }
int adv_plan_frames(const adv_plan* plan) { return plan ? plan->d.T : ADV_ERR_INVALID; }
int adv_plan_tiles(const adv_plan* plan, int batch) {
    if (!plan || batch <= 0 || plan->d.n_out <= 0) return ADV_ERR_INVALID;
    const int s4 = explain4_slots(plan, batch);   // statistics slots per clip of the streaming explain kernels
    if (s4 > 0) return s4;
    const int s5 = explain5_slots(plan, batch);
    return s5 > 0 ? s5 : choose_tiling(plan, batch, 1).tiles;
}
int adv_plan_tiles_istft(const adv_plan* plan, int batch) {
    if (!plan || batch <= 0 || plan->d.n_out <= 0) return ADV_ERR_INVALID;
    const int s4 = istft4_slots(plan, batch);   // statistics slots per clip of the streaming iSTFT
    const int s5 = s4 > 0 ? 0 : istft5_slots(plan, batch);
    if (s5 > 0) return s5;
    return s4 > 0 ? s4 : choose_tiling(plan, batch, 2, istft_balanced(), istft3_frames_cap(plan)).tiles;
}

int adv_stft(const adv_plan* plan, const float* wav, int64_t wav_stride, int batch, adv_c64* X, float* mag,
             float* phase, void* stream) {
    if (!plan || !wav || !X || batch <= 0 || plan->d.n_in <= 0 || wav_stride < plan->d.n_in) return ADV_ERR_INVALID;
    return launch_stft(plan, wav, wav_stride, batch, (float2*)X, mag, phase, 0, (cudaStream_t)stream);
}

int adv_stft_ex(const adv_plan* plan, const float* wav, int64_t wav_stride, int batch, adv_c64* X, float* mag,
                float* phase, int flags, void* stream) {
    if (!plan || !wav || !X || batch <= 0 || plan->d.n_in <= 0 || wav_stride < plan->d.n_in) return ADV_ERR_INVALID;
    if (flags & ~ADV_STFT_ZERO_PAD) return ADV_ERR_INVALID;
    return launch_stft(plan, wav, wav_stride, batch, (float2*)X, mag, phase, flags, (cudaStream_t)stream);
}

int adv_plan_inv_env(const adv_plan* plan, float* out, void* stream) {
    if (!plan || !out || plan->d.n_out <= 0) return ADV_ERR_INVALID;
    ADV_CUDA_CHECK(cudaMemcpyAsync(out, plan->d.inv_env, sizeof(float) * (size_t)plan->d.n_out, cudaMemcpyDeviceToDevice,
                                   (cudaStream_t)stream));
    return ADV_OK;
}

int adv_istft(const adv_plan* plan, const adv_c64* X, int64_t sb, int64_t st, int64_t sf, int batch, float* out,
              double* stats, void* stream) {
    if (!plan || !X || !out || batch <= 0 || plan->d.n_out <= 0) return ADV_ERR_INVALID;
    return launch_istft(plan, (const float2*)X, sb, st, sf, batch, out, stats, (cudaStream_t)stream);
}

int adv_explain(const adv_plan* plan, const float* wav, int64_t wav_stride, const float* mask, int Fm, int Tm, int mode,
                int batch, float* rel, float* irr, double* stats, void* stream) {
    if (!plan || !wav || !mask || !rel || !irr || batch <= 0 || plan->d.n_in <= 0 || wav_stride < plan->d.n_in ||
        plan->d.n_out <= 0)
        return ADV_ERR_INVALID;
    if ((mode & ~ADV_MASK_DROP_OUTSIDE) != ADV_MASK_LOG1P && (mode & ~ADV_MASK_DROP_OUTSIDE) != ADV_MASK_LINEAR) return ADV_ERR_INVALID;
    if (Fm <= 0 || Tm <= 0 || Fm > plan->d.n_fft / 2 + 1 || Tm > plan->d.T) return ADV_ERR_SHAPE;
    return launch_explain(plan, wav, wav_stride, nullptr, 0, 0, 0, mask, Fm, Tm, mode, batch, rel, irr, stats,
                          (cudaStream_t)stream);
}

int adv_explain_spec(const adv_plan* plan, const adv_c64* X, int64_t sb, int64_t st, int64_t sf, const float* mask,
                     int Fm, int Tm, int mode, int batch, float* rel, float* irr, double* stats, void* stream) {
    if (!plan || !X || !mask || !rel || !irr || batch <= 0 || plan->d.n_out <= 0) return ADV_ERR_INVALID;
    if ((mode & ~ADV_MASK_DROP_OUTSIDE) != ADV_MASK_LOG1P && (mode & ~ADV_MASK_DROP_OUTSIDE) != ADV_MASK_LINEAR) return ADV_ERR_INVALID;
    if (Fm <= 0 || Tm <= 0 || Fm > plan->d.n_fft / 2 + 1 || Tm > plan->d.T) return ADV_ERR_SHAPE;
    return launch_explain(plan, nullptr, 0, (const float2*)X, sb, st, sf, mask, Fm, Tm, mode, batch, rel, irr, stats,
                          (cudaStream_t)stream);
}

}  // extern "C"
