// Building-block probe for the tcgen05 path: one CTA computes D[128 x N] = A[128 x K] * B[N x K]^T (bf16 inputs,
// fp32 accumulate in TMEM) through every operand route mpnn_tc.cu uses, so that descriptor / layout mistakes show
// up in isolation (tests/test_gpu_tc_probe.py).  mode bits:
//   bits 0-1  A source : 0 smem K-major, 1 TMEM via tcgen05.st.32x32b, 2 TMEM via tcgen05.st.16x128b
//   bit  2    B layout : 0 K-major, 1 MN-major
//   bit  3    readback : 0 tcgen05.ld.32x32b, 1 tcgen05.ld.16x256b
#include <cuda_bf16.h>

#include "eco_common.cuh"
#include "tc_prims.cuh"

namespace eco {
namespace {

using namespace tc;

__device__ __forceinline__ uint16_t bf16_bits(float x) { return __bfloat16_as_ushort(__float2bfloat16_rn(x)); }

__global__ void __launch_bounds__(128) tc_probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                      float* __restrict__ D, int N, int K, int mode) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int a_mode = mode & 3, b_mn = (mode >> 2) & 1, ld256 = (mode >> 3) & 1;
    unsigned char* sA = smem;                        // 128 x K bf16
    unsigned char* sB = smem + 128 * 208 * 2;         // N x K bf16

    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t TA = 256;                          // TMEM columns for the A operand (K/2 <= 104)

    // ---- operands ------------------------------------------------------------------------------------
    if (a_mode == 0) {
        for (int idx = tid; idx < 128 * K; idx += 128) {
            const int r = idx / K, k = idx % K;
            const int off = ((k >> 3) * 16 + (r >> 3)) * 128 + (r & 7) * 16 + (k & 7) * 2;
            *reinterpret_cast<uint16_t*>(sA + off) = bf16_bits(A[r * K + k]);
        }
    } else if (a_mode == 1) {
        const int r = 32 * warp + lane;               // thread owns TMEM lane r
        for (int ks = 0; ks < K / 16; ++ks) {
            uint32_t v[8];
            for (int c = 0; c < 8; ++c) {
                const int k = ks * 16 + 2 * c;
                v[c] = (uint32_t)bf16_bits(A[r * K + k]) | ((uint32_t)bf16_bits(A[r * K + k + 1]) << 16);
            }
            tmem_st_32x32b_x8(tmem_addr(tmem, 32 * warp, TA + ks * 8), v);
        }
        tmem_st_wait();
    } else {
        for (int half = 0; half < 2; ++half) {
            const int base_row = 32 * warp + 16 * half;
            for (int ks = 0; ks < K / 16; ++ks) {
                uint32_t v[4];
                for (int i = 0; i < 4; ++i) {
                    const int r = base_row + (lane >> 2) + 8 * (i & 1);
                    const int col = (lane & 3) + 4 * (i >> 1);            // packed column within the 8-column group
                    const int k = ks * 16 + 2 * col;
                    v[i] = (uint32_t)bf16_bits(A[r * K + k]) | ((uint32_t)bf16_bits(A[r * K + k + 1]) << 16);
                }
                tmem_st_16x128b_x2(tmem_addr(tmem, base_row, TA + ks * 8), v);
            }
        }
        tmem_st_wait();
    }
    for (int idx = tid; idx < N * K; idx += 128) {
        const int n = idx / K, k = idx % K;
        int off;
        if (!b_mn) off = ((k >> 3) * (N >> 3) + (n >> 3)) * 128 + (n & 7) * 16 + (k & 7) * 2;
        else       off = ((n >> 3) * (K >> 3) + (k >> 3)) * 128 + (k & 7) * 16 + (n & 7) * 2;
        *reinterpret_cast<uint16_t*>(sB + off) = bf16_bits(B[n * K + k]);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();

    // ---- MMA -------------------------------------------------------------------------------------------
    if (tid == 0) {
        tc_fence_after();
        const uint32_t idesc = instr_desc_bf16(128, N, false, b_mn != 0);
        for (int ks = 0; ks < K / 16; ++ks) {
            uint64_t bd;
            if (!b_mn) bd = smem_desc(smem_u32(sB) + ks * 2 * (N >> 3) * 128, (N >> 3) * 128, 128);
            else       bd = smem_desc(smem_u32(sB) + ks * 2 * 128, 128, (K >> 3) * 128);
            if (a_mode == 0) {
                const uint64_t ad = smem_desc(smem_u32(sA) + ks * 2 * 2048, 2048, 128);
                mma_ss(tmem, ad, bd, idesc, ks > 0);
            } else {
                mma_ts(tmem, tmem + TA + ks * 8, bd, idesc, ks > 0);
            }
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();

    // ---- read back ---------------------------------------------------------------------------------------
    if (!ld256) {
        const int r = 32 * warp + lane;
        for (int c0 = 0; c0 < N; c0 += 8) {
            uint32_t v[8];
            tmem_ld_32x32b_x8(tmem_addr(tmem, 32 * warp, c0), v);
            tmem_ld_wait();
            for (int c = 0; c < 8; ++c) D[r * N + c0 + c] = __uint_as_float(v[c]);
        }
    } else {
        for (int half = 0; half < 2; ++half) {
            const int base_row = 32 * warp + 16 * half;
            for (int c0 = 0; c0 < N; c0 += 16) {
                uint32_t v[8];
                tmem_ld_16x256b_x2(tmem_addr(tmem, base_row, c0), v);
                tmem_ld_wait();
                for (int i = 0; i < 8; ++i) {
                    const int r = base_row + (lane >> 2) + 8 * ((i >> 1) & 1);
                    const int c = c0 + 8 * (i >> 2) + 2 * (lane & 3) + (i & 1);
                    D[r * N + c] = __uint_as_float(v[i]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace
}  // namespace eco

// Debug entry point (not part of the public header): A [128,K], B [N,K], D [128,N] fp32 device pointers.
extern "C" int eco_tc_probe(const float* A_dev, const float* B_dev, float* D_dev, int N, int K, int mode, void* stream) {
    using namespace eco;
    ECO_CHECK_ARG(A_dev && B_dev && D_dev, ECO_ERR_INVALID, "eco_tc_probe: null argument");
    ECO_CHECK_ARG(N % 16 == 0 && N >= 16 && N <= 208 && K % 16 == 0 && K >= 16 && K <= 208, ECO_ERR_INVALID,
                  "eco_tc_probe: need N, K multiples of 16 in [16, 208]");
    const int smem_bytes = 128 * 208 * 2 + 208 * 208 * 2;
    ECO_CUDA(cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    tc_probe_kernel<<<1, 128, smem_bytes, (cudaStream_t)stream>>>(A_dev, B_dev, D_dev, N, K, mode);
    ECO_LAUNCH_CHECK();
    return ECO_OK;
}
