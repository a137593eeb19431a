// Micro-benchmark: how many tcgen05.mma instructions can be in flight before the issuing thread blocks?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_prims.cuh"
using namespace eco::tc;

__global__ void __launch_bounds__(128, 1) k(int kmma, int N, long long* out) {
    __shared__ __align__(128) unsigned char sa[4096];
    extern __shared__ __align__(128) unsigned char sb[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 4096 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sa)[i] = 0;
    for (int i = threadIdx.x; i < 256 * 32 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sb)[i] = 0;
    if (warp == 0) tmem_alloc(&tbase, 512);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tbase;
    if (warp == 0) {
        long long t0 = 0, t1 = 0, t2 = 0;
        if (elect_one()) {
            const uint32_t idesc = instr_desc_bf16(128, N, false, false);
            const uint64_t ad = smem_desc(smem_u32(sa), 2048, 128);
            const uint64_t bd = smem_desc(smem_u32(sb), (N / 8) * 128, 128);
            t0 = clock64();
            for (int i = 0; i < kmma; ++i) mma_ss(tmem + 256, ad, bd, idesc, i > 0);
            t1 = clock64();
            mma_commit(&bar);
            t2 = clock64();
            out[0] = t1 - t0;
            out[1] = t2 - t1;
        }
        __syncwarp();
        long long t3 = clock64();
        mbar_wait(&bar, 0);
        long long t4 = clock64();
        if (lane == 0) { out[2] = t4 - t3; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    long long* d;
    cudaMalloc(&d, 32 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 32);
    for (int N : {64, 208}) {
        for (int kk : {1, 2, 4, 8, 12, 16, 24, 32, 48, 64, 128}) {
            long long h[3];
            for (int rep = 0; rep < 2; ++rep) {
                cudaMemset(d, 0, 32 * 8);
                k<<<1, 128, 256 * 32>>>(kk, N, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
            }
            printf("N=%3d  %3d MMAs: issue loop %6lld cyc, commit %4lld cyc, wait after %6lld cyc\n", N, kk, h[0], h[1], h[2]);
        }
    }
    return 0;
}
