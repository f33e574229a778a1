#include <cstdio>
#include <cuda_runtime.h>
template <int COLS, bool IMM>
__global__ void __launch_bounds__(384, 2) k3(float* o) {
    extern __shared__ float s[];
    unsigned* slot = reinterpret_cast<unsigned*>(s + 1024);
    if (threadIdx.x < 32) {
        if (IMM) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(slot)), "n"(COLS) : "memory");
        else asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(slot)), "r"((unsigned)COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    o[threadIdx.x] = s[threadIdx.x ^ 1] + *slot;
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*slot), "r"((unsigned)COLS) : "memory");
}
template <class K> void probe(K kern, const char* name) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    printf("%s: regs %d:", name, fa.numRegs);
    for (int kb = 16; kb <= 112; kb += 32) { int n = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, 384, kb * 1024); printf("  %d KB -> %d blocks;", kb, n); }
    printf("\n");
    // real residency: launch 296 blocks that spin until all have started (max 2 s)
}
__global__ void __launch_bounds__(384, 2) spin(int* counter, int target, int* ok) {
    extern __shared__ float s[];
    unsigned* slot = reinterpret_cast<unsigned*>(s + 1024);
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(slot)), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(counter, 1);
        long long t0 = clock64();
        while (atomicAdd(counter, 0) < target && clock64() - t0 < 2000000000LL) {}
        if (atomicAdd(counter, 0) >= target) atomicAdd(ok, 1); else atomicAdd(ok + 1, 1);   // ok[0] = saw all, ok[1] = timed out
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*slot), "n"(256) : "memory");
}
int main() {
    probe(k3<256, false>, "tmem256 reg "); probe(k3<256, true>, "tmem256 imm "); probe(k3<128, true>, "tmem128 imm "); probe(k3<512, true>, "tmem512 imm ");
    int *c, *ok; cudaMalloc(&c, 4); cudaMalloc(&ok, 8); cudaMemset(c, 0, 4); cudaMemset(ok, 0, 8);
    cudaFuncSetAttribute(spin, cudaFuncAttributeMaxDynamicSharedMemorySize, 108 * 1024);
    cudaFuncSetAttribute(spin, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    spin<<<296, 384, 108 * 1024>>>(c, 296, ok);
    cudaError_t e = cudaDeviceSynchronize();
    int h[2] = {0, 0}; cudaMemcpy(h, ok, 8, cudaMemcpyDeviceToHost);
    printf("co-residency test (296 CTAs of 108 KB, 256 TMEM cols each): %d saw all 296 alive, %d timed out -> %s (%s)\n", h[0], h[1], h[1] == 0 ? "YES, 2 CTAs/SM" : "NO", cudaGetErrorString(e));
    return 0;
}
