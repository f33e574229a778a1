// Micro-benchmark: tcgen05.ld (TMEM -> registers) throughput per SM as a function of the number of warps.
// Each warp reads 32 lanes x 32 columns (4 KB) per instruction from its own lane quadrant (warp % 4).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_rate tmem_ld_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void k(float* out, int iters, int mode) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t r[32];
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t addr = base + ((it * 32 + (warp >> 2) * 64) & 511 & ~31);
        if (mode == 0) {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(addr) : "memory");
        } else {   // 16-bit packed view: two 16-bit columns per register (what an fp16 accumulator would need)
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(addr) : "memory");
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc += r[0] ^ r[15];
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

int main() {
    float* d; cudaMalloc(&d, 148 * 1024 * sizeof(float));
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode)
        for (int warps : {1, 2, 4, 8, 16}) {
            k<<<148, warps * 32>>>(d, 100, mode);
            k<<<148, warps * 32>>>(d, iters, mode);
            cudaError_t e = cudaDeviceSynchronize();
            float cyc; cudaMemcpy(&cyc, d, 4, cudaMemcpyDeviceToHost);
            const double bytes = (double)warps * iters * (mode == 0 ? 4096.0 : 2048.0);
            printf("%s  %2d warps/SM: %7.1f bytes/clk/SM  (%.0f cycles per ld per warp)  %s\n", mode == 0 ? "x32" : "x16", warps,
                   bytes / cyc, cyc / iters, cudaGetErrorString(e));
        }
    return 0;
}
