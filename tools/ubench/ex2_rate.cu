// Micro-benchmark: MUFU.EX2 throughput per SM for f32 vs packed f16x2 / bf16x2 arguments, and a degree-3 polynomial
// exp2 on the FMA pipe.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ex2_rate ex2_rate.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
    float a[8];
    uint32_t h[8];
    for (int i = 0; i < 8; ++i) { a[i] = seed * (threadIdx.x + i) * 1e-3f; h[i] = 0x3c003800u + threadIdx.x + i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); }
            if (MODE == 1) { asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i])); }
            if (MODE == 2) { asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i])); }
            if (MODE == 3) {   // Cody-Waite + degree-3 polynomial on the FMA/ALU pipes
                float x = a[i];
                float fl = floorf(x);
                float f = x - fl;
                float p = fmaf(f, 0.0555041f, 0.2402265f);
                p = fmaf(p, f, 0.6931472f);
                p = fmaf(p, f, 1.0f);
                a[i] = __int_as_float(__float_as_int(p) + ((int)fl << 23)) * 1e-3f;
            }
            if (MODE == 5) {   // one ex2 and four dependent-free FFMA per element: the softmax mix
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                float t = __uint_as_float(h[i]);
                t = fmaf(t, 1.0001f, 0.5f); t = fmaf(t, 0.9999f, -0.5f); t = fmaf(t, 1.0001f, 0.25f); t = fmaf(t, 0.9999f, -0.25f);
                h[i] = __float_as_uint(t);
            }
            if (MODE == 6) {   // FFMA only, 4 per element
                float t = __uint_as_float(h[i]);
                t = fmaf(t, 1.0001f, 0.5f); t = fmaf(t, 0.9999f, -0.5f); t = fmaf(t, 1.0001f, 0.25f); t = fmaf(t, 0.9999f, -0.25f);
                h[i] = __float_as_uint(t);
            }
            if (MODE == 7) {   // ex2 + 2 FFMA + 2 FMNMX (fma pipe + alu pipe)
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                float t = __uint_as_float(h[i]);
                t = fmaf(t, 1.0001f, 0.5f); t = fmaxf(t, a[(i + 1) & 7]); t = fmaf(t, 0.9999f, -0.5f); t = fminf(t, 3.0f);
                h[i] = __float_as_uint(t);
            }
            if (MODE == 8) {   // ex2 + 1 FFMA + 1 FADD + half F2FP: the minimal softmax mix
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                float t = __uint_as_float(h[i]);
                t = fmaf(t, 1.0001f, 0.5f); t = t + a[(i + 3) & 7];
                h[i] = __float_as_uint(t);
            }
            if (MODE == 9) {   // ex2 + half a cvt.rn.bf16x2.f32 per element (the P pack)
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                if (i & 1) { uint32_t pk; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(a[i]), "f"(a[i - 1])); h[i] ^= pk; }
            }
            if (MODE == 10) {  // cvt.rn.bf16x2.f32 only
                uint32_t pk; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(a[i]), "f"(a[(i + 1) & 7])); h[i] ^= pk;
            }
            if (MODE == 11) {  // full softmax mix: FFMA, ex2, FADD, half cvt, half FMNMX3-like
                float x = fmaf(__uint_as_float(h[i]), 1.0001f, a[(i + 1) & 7]);
                float p; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(x));
                a[i] += p;
                if (i & 1) { uint32_t pk; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(p), "f"(x)); h[i] ^= pk; }
            }
            if (MODE == 4) {   // cvt pair + packed ex2 (the softmax inner step)
                uint32_t pk;
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(a[i]), "f"(a[(i + 1) & 7]));
                asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(pk));
                h[i] ^= pk;
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += a[i] + (float)h[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}

template <int MODE>
void run(const char* name, int elems_per_instr, int threads = 1024) {
    float* d; cudaMalloc(&d, 148 * 1024 * sizeof(float));
    const int iters = 4096;
    k<MODE><<<148, threads>>>(d, 16, 1.0f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<148, threads>>>(d, iters, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    float cyc; cudaMemcpy(&cyc, d, 4, cudaMemcpyDeviceToHost);
    double el = (double)threads * iters * 8 * elems_per_instr;
    printf("[%4d thr] %-28s %8.3f ms  %7.2f elements/clk/SM (clock64)  %.2f T elem/s chip\n", threads, name, ms, el / cyc,
           el * 148 / ms / 1e9);
    cudaFree(d);
}

int main() {
    run<0>("ex2.approx.ftz.f32", 1);
    run<0>("ex2.approx.ftz.f32", 1, 512);
    run<0>("ex2.approx.ftz.f32", 1, 256);
    run<0>("ex2.approx.ftz.f32", 1, 128);
    run<5>("ex2 + 4 FFMA interleaved", 1, 512);
    run<5>("ex2 + 4 FFMA interleaved", 1, 1024);
    run<6>("4 FFMA (no ex2)", 1, 512);
    run<9>("ex2 + 0.5 cvt.bf16x2", 1, 512);
    run<10>("cvt.bf16x2 only (per instr)", 1, 512);
    run<11>("FFMA+ex2+FADD+0.5cvt", 1, 512);
    run<7>("ex2 + 2 FFMA + 2 FMNMX", 1, 512);
    run<8>("ex2 + FFMA + FADD", 1, 512);
    run<1>("ex2.approx.f16x2", 2);
    run<2>("ex2.approx.ftz.bf16x2", 2);
    run<3>("poly3 exp2 (FMA pipe)", 1);
    run<4>("cvt.f16x2 + ex2.f16x2", 2);
    return 0;
}
