// Packed bf16 hi/lo weight operands (eco_mpnn_pack output) shared by the tensor-core kernels.
#pragma once
#include <stdint.h>

namespace eco {

// packed weights (uint32 words): [128 stacked rows][k/2] per matrix
constexpr int PK_WEF = 0;                    // 128 x 32
constexpr int PK_WM = PK_WEF + 128 * 32;     // 3 x 128 x 64
constexpr int PK_WU = PK_WM + 3 * 128 * 64;  // 3 x 128 x 64
constexpr int PK_WPT = PK_WU + 3 * 128 * 64;  // 64 x 64 fp32: W_p transposed ([k][f]), read coalesced by the one-warp readout
constexpr int PK_WORDS = PK_WPT + 64 * 64;

// Packed layout per matrix (KW words per stacked row): word (r, c) with r = 32q + lane, c = 8cg + 4h + j lives at
// ((((q * KW/8 + cg) * 2 + h) * 32 + lane) * 4 + j): every LDG.128 of a warp in ldg_weights() is 512 contiguous bytes.
__host__ __device__ inline int packed_index(int r, int c, int kw) {
    const int q = r >> 5, lane = r & 31, cg = c >> 3, h = (c >> 2) & 1, j = c & 3;
    return (((q * (kw / 8) + cg) * 2 + h) * 32 + lane) * 4 + j;
}


}  // namespace eco
