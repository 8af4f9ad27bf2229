// Micro-benchmark of per-SM instruction throughput on sm_100a: scalar vs packed fp32 (FFMA / FFMA2), MUFU.EX2 in fp32 and
// f16x2, and their mixes.  One CTA of 1024 threads per SM, ITER dependent-chain steps over 8 independent chains per thread.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int ITER = 4096;

template <int MODE>
__global__ void k(float* out, float seed) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-6f + i;
    unsigned long long p[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
    unsigned h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = 0x3c003800u + i;
    const float c = seed * 0.999f;
    unsigned long long c2;
    asm("mov.b64 %0, {%1, %1};" : "=l"(c2) : "f"(c));
    for (int it = 0; it < ITER; ++it) {
        if (MODE == 0) {          // 8 FFMA
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], c, c);
        } else if (MODE == 1) {   // 4 FFMA2 (same flops as mode 0)
#pragma unroll
            for (int i = 0; i < 4; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(c2));
        } else if (MODE == 2) {   // 8 MUFU.EX2 fp32
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        } else if (MODE == 3) {   // 8 MUFU.EX2 f16x2 (16 exps)
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
        } else if (MODE == 4) {   // 8 FFMA2 (twice the flops of mode 0)
#pragma unroll
            for (int i = 0; i < 4; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(c2));
#pragma unroll
            for (int i = 0; i < 4; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(c2));
        } else if (MODE == 5) {   // 2 MUFU + 8 FFMA interleaved (is the issue slot shared?)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], c, c);
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[0]));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[4]));
        } else if (MODE == 6) {   // 2 MUFU + 16 FFMA
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], c, c);
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], c, c);
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[0]));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[4]));
        } else if (MODE == 7) {   // 2 MUFU + 8 FFMA2
#pragma unroll
            for (int i = 0; i < 4; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(c2));
#pragma unroll
            for (int i = 0; i < 4; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(c2));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[0]));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[4]));
        } else if (MODE == 8) {   // 8 MUFU.RCP
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        } else if (MODE == 9) {   // 8 LEA-like integer ops (ALU pipe) + 8 FFMA
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], c, c);
#pragma unroll
            for (int i = 0; i < 8; ++i) h[i] = (h[i] << 3) + h[(i + 1) & 7];
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p[i])); s += x + y; }
    if (s == 12345.678f) out[0] = s;
}

template <int MODE>
void run(const char* name, int ops_per_iter, float* d) {
    int dev = 0, sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms, 1024>>>(d, 1.0f);
    cudaEventRecord(e0);
    k<MODE><<<sms, 1024>>>(d, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * khz * 1e3;                       // cycles at the max clock
    const double warp_instr = 32.0 * ITER * ops_per_iter;           // per SM
    printf("%-34s %8.3f ms  %6.2f cycles per warp instruction per SM (%5.2f per scheduler)\n", name, ms, cyc / warp_instr,
           4.0 * cyc / warp_instr);
}

int main() {
    float* d; cudaMalloc(&d, 4);
    run<0>("8 FFMA", 8, d);
    run<1>("4 FFMA2", 4, d);
    run<4>("8 FFMA2", 8, d);
    run<2>("8 MUFU.EX2 f32", 8, d);
    run<3>("8 MUFU.EX2 f16x2", 8, d);
    run<8>("8 MUFU.RCP", 8, d);
    run<5>("2 MUFU + 8 FFMA", 10, d);
    run<6>("2 MUFU + 16 FFMA", 18, d);
    run<7>("2 MUFU + 8 FFMA2", 10, d);
    run<9>("8 FFMA + 8 shift-add", 16, d);
    return 0;
}
