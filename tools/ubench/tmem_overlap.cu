// Micro-benchmark: does tcgen05.ld throughput drop while tcgen05.mma runs?  (tools only, not part of the library)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I eco-dqn_b200/csrc -o tools/ubench/tmem_overlap tools/ubench/tmem_overlap.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_prims.cuh"
using namespace eco::tc;

__global__ void __launch_bounds__(544, 1) k(int mode, int nld, int nmma, int N, int nwarps_ld, int x1, long long* out) {
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
    long long t0 = clock64();
    if (warp == 16) {
        if (mode & 1) {
            if (elect_one()) {
                const uint32_t idesc = instr_desc_bf16(128, N, false, false);
                const uint64_t ad = smem_desc(smem_u32(sa), 2048, 128);
                const uint64_t bd = smem_desc(smem_u32(sb), (N / 8) * 128, 128);
                for (int i = 0; i < nmma; ++i) mma_ss(tmem + 256, ad, bd, idesc, i > 0);
                mma_commit(&bar);
            }
            __syncwarp();
            mbar_wait(&bar, 0);
            long long t1 = clock64();
            if (lane == 0) out[16] = t1 - t0;
        }
    } else if (warp < nwarps_ld && (mode & 2)) {
        const int q = warp & 3, sub = warp >> 2;
        float acc = 0.f;
        for (int i = 0; i < nld; ++i) {
            uint32_t a[8], b[8];
            const int col = ((i * 4 + sub) * 16) & 255;
            if (x1) {
                tmem_ld_32x32b_x8(tmem_addr(tmem, 32 * q, col), a);
                tmem_ld_32x32b_x8(tmem_addr(tmem, 32 * q, col + 8), b);
            } else {
                tmem_ld_16x256b_x2(tmem_addr(tmem, 32 * q, col), a);
                tmem_ld_16x256b_x2(tmem_addr(tmem, 32 * q + 16, col), b);
            }
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += __uint_as_float(a[j]) + __uint_as_float(b[j]);
        }
        long long t1 = clock64();
        if (lane == 0) out[warp] = t1 - t0;
        if (acc == 123.f) out[20] = 1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    long long* d;
    cudaMalloc(&d, 32 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 32);
    
    for (int x1 = 0; x1 < 2; ++x1)
    for (int nw : {4, 8, 12, 16}) {
        for (int N : {48, 64, 208}) {
            const int nmma = 8192 * 64 / N; const int nld = 8192 / nw;
            long long h[3][32];
            for (int mode = 1; mode <= 3; ++mode) {
                cudaMemset(d, 0, 32 * 8);
                k<<<1, 544, 256 * 32>>>(mode, nld, nmma, N, nw, x1, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h[mode - 1], d, 32 * 8, cudaMemcpyDeviceToHost);
            }
            long long ld_alone = 0, ld_both = 0;
            for (int w = 0; w < nw; ++w) { if (h[1][w] > ld_alone) ld_alone = h[1][w]; if (h[2][w] > ld_both) ld_both = h[2][w]; }
            const double bytes = (double)nld * nw * 2048;
            printf("%s ld warps %2d  N=%3d: mma alone %7lld cyc (%.1f cyc/mma)  ld alone %7lld cyc (%.1f B/clk) | together: mma %7lld (%.1f cyc/mma)  ld %7lld (%.1f B/clk)\n",
                   x1 ? "32x32b.x8 " : "16x256b.x2", nw, N, h[0][16], (double)h[0][16] / nmma, ld_alone, bytes / ld_alone, h[2][16],
                   (double)h[2][16] / nmma, ld_both, bytes / ld_both);
        }
    }
    return 0;
}
