// throughput of FADD/FFMA vs packed FADD2/FFMA2 on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
    float a0 = threadIdx.x, a1 = 1.f, a2 = 2.f, a3 = 3.f, a4 = 4.f, a5 = 5.f, a6 = 6.f, a7 = 7.f;
    float b = 1.0001f, c = 0.5f;
    u64 p0, p1, p2, p3, pb, pc;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(a2), "f"(a3));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(a4), "f"(a5));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(a6), "f"(a7));
    asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b), "f"(b));
    asm("mov.b64 %0, {%1, %2};" : "=l"(pc) : "f"(c), "f"(c));
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (MODE == 0) {  // 8 scalar FFMA
                a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
                a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
            } else if (MODE == 1) {  // 4 packed FFMA2 (same flops)
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pb), "l"(pc));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(pb), "l"(pc));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(pb), "l"(pc));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(pb), "l"(pc));
            } else if (MODE == 2) {  // 8 scalar FADD
                a0 += c; a1 += c; a2 += c; a3 += c; a4 += c; a5 += c; a6 += c; a7 += c;
            } else {  // 4 packed FADD2
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p0) : "l"(pc));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p1) : "l"(pc));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p2) : "l"(pc));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p3) : "l"(pc));
            }
        }
    }
    float r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    float x, y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p0)); r += x + y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p1)); r += x + y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p2)); r += x + y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p3)); r += x + y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, float* d) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 4096, blocks = 148 * 8;
    k<MODE><<<blocks, 256>>>(d, 16);
    cudaEventRecord(a);
    k<MODE><<<blocks, 256>>>(d, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double flops = 2.0 * 8 * 16 * (double)iters * blocks * 256 * ((MODE >= 2) ? 0.5 : 1.0);
    printf("%s: %.3f ms, %.1f T(fl)op/s fp32-lane-ops\n", name, ms, flops / ms / 1e9);
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("FFMA  x8", d); run<1>("FFMA2 x4", d); run<2>("FADD  x8", d); run<3>("FADD2 x4", d);
    return 0;
}
