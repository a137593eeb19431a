// Kernel family 2c: tensor-core neighbour aggregation for graphs that do not fit mpnn_tc_kernel (208 < N <= 2048,
// couplings in {-1,0,1}).
//
// Replaces (reference, file:line) the two N x N products of the MPNN:
//   src/networks/mpnn.py:114-116  UpdateNodeEmbeddingLayer: torch.matmul(adj, node_features) / norm
//   src/networks/mpnn.py:89-102   EdgeAndNodeEmbeddingLayer: sum_j ReLU(W_e [a_ij ; x_j]) [a_ij != 0] / norm, in the
//                                 factorised form 1/2 (|A| S + A D) of mpnn_tc.cu (S = R+ + R-, D = R+ - R-)
// The per-vertex linears around them stay on the CUDA-core kernel (mpnn_simt.cu, phase mode): at N = 500, p = 0.15 the
// sparse neighbour visit is ~75 % of that kernel, while the dense contraction is 64 MFLOP per layer and episode.
//
//   OUT[b][i][f] = scale / deg_i * sum_j ( X1[b][j][f] * IMG1_g[j][i]  (+ X2[b][j][f] * IMG2_g[j][i]) )
//
// One CTA per (episode, 128-column slab).  K is walked in 64-vertex panels: the fp32 rows X[j][0..63] are split into
// bf16 hi/lo, stacked along M (128 rows, same row order as mpnn_tc.cu) and written as a K-major A-operand; the matching
// panel of the graph's bf16 operand image (graph_prepare.cu: tc_ops) arrives by bulk copies; tcgen05.mma accumulates the
// slab in TMEM (128 columns, so four CTAs share an SM and overlap each other's load / convert / MMA phases -- there is
// no pipeline inside a CTA).
#include <cuda_bf16.h>

#include "eco_common.cuh"
#include "tc_prims.cuh"
#include "mpnn_pack.cuh"

namespace eco {
namespace {

using namespace tc;

constexpr int SLAB = 128;        // accumulator columns (vertices i) per CTA
constexpr int PANEL = 64;        // vertices j per K panel
constexpr int TCL_THREADS = 128;
constexpr int XF_BYTES = 64 * 65 * 4 + 256;                 // fp32 staging [64][65], padded to a multiple of 128 bytes
static_assert(XF_BYTES % 128 == 0, "staging size");
constexpr int OP_BYTES = 128 * PANEL * 2;                   // one operand panel (A: 128 rows x 64 k; B: 64 k x 128 cols)

template <int PAIRS>
__global__ void __launch_bounds__(TCL_THREADS)
tcl_contract_kernel(const eco_graphs_t g, const int32_t* __restrict__ graph_idx, const int B,
                    const float* __restrict__ X1, const int which1, const float* __restrict__ X2, const int which2,
                    const size_t x_stride, float* __restrict__ out, const size_t out_stride, const float scale,
                    const int edge, const float norm_max) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar_b, bar_mma;
    __shared__ uint32_t tmem_base_s;
    float* sXf = reinterpret_cast<float*>(smem);
    unsigned char* sX[2] = {smem + XF_BYTES, smem + XF_BYTES + 2 * OP_BYTES};
    unsigned char* sB[2] = {smem + XF_BYTES + OP_BYTES, smem + XF_BYTES + 3 * OP_BYTES};
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int N = g.N, NP = g.NP, NB = NP >> 3;
    const int nslabs = (NP + SLAB - 1) / SLAB;
    const int b = blockIdx.x / nslabs, slab = blockIdx.x % nslabs;
    const int n0 = slab * SLAB, w = min(SLAB, NP - n0);      // NP is a multiple of 16, so is w
    const int gi = graph_idx[b];
    const float* Xp[2] = {X1 + (size_t)b * x_stride, PAIRS == 2 ? X2 + (size_t)b * x_stride : nullptr};
    const uint16_t* img[2] = {g.tc_ops + ((size_t)gi * 2 + which1) * NP * NP,
                              g.tc_ops + ((size_t)gi * 2 + (PAIRS == 2 ? which2 : which1)) * NP * NP};

    if (warp == 0) tmem_alloc(&tmem_base_s, 128);
    if (tid == 0) { mbar_init(&bar_b, 1); mbar_init(&bar_mma, 1); fence_mbar_init(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    uint32_t phase_b = 0, phase_m = 0;
    // this thread's stacked operand row: r = 32q + 16s + t  <->  feature 16q + t, split s (hi / lo)
    const int r = tid, f = 16 * (r >> 5) + (r & 15), split = (r >> 4) & 1;
    const int npanels = (NP + PANEL - 1) / PANEL;
    const int run = (w >> 3) * 128;                            // bytes of one 8-vertex K group of the B panel

    for (int kp = 0; kp < npanels; ++kp) {
        const int k0 = kp * PANEL, kw = min(PANEL, NP - k0);
        if (kp > 0) { mbar_wait(&bar_mma, phase_m); phase_m ^= 1u; }       // the previous panel's MMAs have read smem
        if (tid == 0) {
            mbar_expect_tx(&bar_b, (uint32_t)(PAIRS * (kw >> 3) * run));
            for (int p = 0; p < PAIRS; ++p)
                for (int cb = 0; cb < (kw >> 3); ++cb)
                    bulk_g2s(sB[p] + cb * run, img[p] + ((size_t)((k0 >> 3) + cb) * NB + (n0 >> 3)) * 64, run, &bar_b);
        }
        for (int p = 0; p < PAIRS; ++p) {
            if (p > 0) __syncthreads();                                     // sXf is reused
            for (int idx = tid; idx < kw * 64; idx += TCL_THREADS) {        // coalesced fp32 rows of the panel
                const int jj = idx >> 6, ff = idx & 63;
                sXf[jj * 65 + ff] = (k0 + jj < N) ? Xp[p][(size_t)(k0 + jj) * 64 + ff] : 0.f;
            }
            __syncthreads();
            for (int jj = 0; jj < kw; jj += 2) {
                uint32_t hi, lo;
                split2(sXf[jj * 65 + f], sXf[(jj + 1) * 65 + f], hi, lo);
                *reinterpret_cast<uint32_t*>(sX[p] + (((jj >> 3) * 16 + (r >> 3)) * 128 + (r & 7) * 16 + (jj & 7) * 2)) =
                    split ? lo : hi;
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            mbar_wait(&bar_b, phase_b);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t idesc = instr_desc_bf16(128, w, false, false);
                for (int p = 0; p < PAIRS; ++p) {
                    const uint64_t ad = smem_desc(smem_u32(sX[p]), 2048, 128);
                    const uint64_t bd = smem_desc(smem_u32(sB[p]), run, 128);
                    for (int ks = 0; ks < (kw >> 4); ++ks)
                        mma_ss(tmem, ad + (uint64_t)ks * (4096 >> 4), bd + (uint64_t)ks * ((2 * run) >> 4), idesc,
                               kp > 0 || p > 0 || ks > 0);
                }
                mma_commit(&bar_mma);
            }
            __syncwarp();
        }
        phase_b ^= 1u;
    }
    mbar_wait(&bar_mma, phase_m);
    tc_fence_after();

    // epilogue: hi + lo rows, scale, 1 / deg; feature 63 of the edge stage = deg / deg_max (mpnn.py:102)
    const float dmax = norm_max > 0.f ? norm_max : (norm_max < 0.f ? (float)max(g.gstat[(size_t)gi * 4], 1) : *g.dmax);
    float* ob = out + (size_t)b * out_stride;
    for (int blk = 0; blk < (w >> 4); ++blk) {
        uint32_t vh[8], vl[8];
        tmem_ld_16x256b_x2(tmem_addr(tmem, 32 * warp, 16 * blk), vh);
        tmem_ld_16x256b_x2(tmem_addr(tmem, 32 * warp + 16, 16 * blk), vl);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int n = n0 + 16 * blk + 8 * (i >> 2) + 2 * (lane & 3) + (i & 1);
            const int ff = 16 * warp + (lane >> 2) + 8 * ((i >> 1) & 1);
            if (n < N) {
                const float d = g.deg[(size_t)gi * NP + n];
                const float v = (__uint_as_float(vh[i]) + __uint_as_float(vl[i])) * scale;
                ob[(size_t)n * 64 + ff] = (edge && ff == 63) ? d / dmax : v / d;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

// ------------------------------------------------------------------------------------------------ per-vertex linears
// The per-vertex linears of a layer on the tensor cores, for 64-vertex slabs of [*][64] fp32 rows in global memory:
//   MODE 0 (edge features):  E  = ReLU(W_ef G)                                   (mpnn.py:100-104)
//   MODE 1 (layer l):        m  = ReLU(W_m [AGG ; E]),  H' = ReLU(W_u [H ; m])    (mpnn.py:117-120)
// Same formulation as mpnn_tc.cu: weights as A-operands in TMEM (bf16 hi/lo rows stacked along M), activations split into
// bf16 hi/lo and stacked (MN-major B operands), two MMA chains per 64 input features, hi + lo rows added in the epilogue.
// Persistent CTAs (weights are loaded into TMEM once per CTA), two per SM (256 TMEM columns each).
constexpr int LSLAB = 64;                                   // vertices per work item
constexpr uint32_t TL_WA = 0, TL_WB = 64, TL_ACCM = 128, TL_ACCH = 192;
constexpr int LXF = 64 * 65 * 4 + 256;                     // one fp32 staging tile
constexpr int LOP = 128 * LSLAB * 2;                       // one stacked operand tile
constexpr int LSMEM = 3 * LXF + 3 * LOP;

// 8 MMAs: acc (+)= W[:, 64-feature group at TMEM column tw] * X, X a stacked operand tile (MN-major)
__device__ __forceinline__ void issue_linear_half(uint32_t tmem, uint32_t acc_col, uint32_t tw, const unsigned char* x, int width,
                                                  bool accumulate) {
    const uint32_t idesc = instr_desc_bf16(128, width, false, true);
    const uint64_t d = smem_desc(smem_u32(x), /*LBO (k groups)*/ 128, /*SBO (vertex groups)*/ 2048);
#pragma unroll
    for (int i = 0; i < 8; ++i)     // i = 2*kq + s: features 16kq..16kq+15, split s
        mma_ts(tmem + acc_col, tmem + tw + 8 * (i >> 1), d + (uint64_t)(16 * i), idesc, accumulate || i > 0);
}

template <int MODE>
__global__ void __launch_bounds__(128, 2)
tcl_linear_kernel(const eco_graphs_t g, const eco_mpnn_t w, const int B, float* __restrict__ buf, const int layer) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    float* sXf[3] = {reinterpret_cast<float*>(smem), reinterpret_cast<float*>(smem + LXF), reinterpret_cast<float*>(smem + 2 * LXF)};
    unsigned char* sOp[3] = {smem + 3 * LXF, smem + 3 * LXF + LOP, smem + 3 * LXF + 2 * LOP};
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int N = g.N, NP = g.NP;
    const size_t bs = (size_t)NP * 64, es = 6 * bs;
    const int nslabs = (NP + LSLAB - 1) / LSLAB;
    const uint32_t* pk = reinterpret_cast<const uint32_t*>(w.packed);

    if (warp == 0) tmem_alloc(&tmem_base_s, 256);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    uint32_t phase = 0;
    {   // this layer's weights: packed global -> registers -> TMEM (warp q owns lane quadrant q)
        auto load = [&](const uint32_t* m, int kw, uint32_t tcol) {
            const uint4* src = reinterpret_cast<const uint4*>(m) + (size_t)(warp * (kw / 8) * 2) * 32 + lane;
            for (int cg = 0; cg < kw / 8; ++cg) {
                const uint4 x = __ldg(src + (2 * cg) * 32), y = __ldg(src + (2 * cg + 1) * 32);
                const uint32_t v[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
                tmem_st_32x32b_x8(tmem_addr(tmem, 32 * warp, tcol + cg * 8), v);
            }
        };
        if (MODE == 0) load(pk + PK_WEF, 32, TL_WA);
        else { load(pk + PK_WM + layer * 128 * 64, 64, TL_WA); load(pk + PK_WU + layer * 128 * 64, 64, TL_WB); }
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
    }
    // stacked operand row of this thread (conversion) and its epilogue coordinates
    const int r = tid, f = 16 * (r >> 5) + (r & 15), split = (r >> 4) & 1;
    const int in0 = 3, in1 = 2, in2 = (layer & 1) ? 1 : 0, outb = MODE == 0 ? 2 : ((layer & 1) ? 0 : 1);
    const int nin = MODE == 0 ? 1 : 3;

    for (int item = blockIdx.x; item < B * nslabs; item += gridDim.x) {
        const int b = item / nslabs, n0 = (item % nslabs) * LSLAB, wdt = min(LSLAB, NP - n0);
        float* eb = buf + (size_t)b * es;
        const int srcs[3] = {in0, in1, in2};
        for (int p = 0; p < nin; ++p) {                   // AGG (or G), E, H rows of the slab: coalesced fp32 -> staging
            const float* X = eb + (size_t)srcs[p] * bs + (size_t)n0 * 64;
            for (int idx = tid; idx < wdt * 64; idx += 128)
                sXf[p][(idx >> 6) * 65 + (idx & 63)] = (n0 + (idx >> 6) < N) ? X[idx] : 0.f;
        }
        __syncthreads();
        for (int p = 0; p < nin; ++p)
            for (int jj = 0; jj < wdt; jj += 2) {
                uint32_t hi, lo;
                split2(sXf[p][jj * 65 + f], sXf[p][(jj + 1) * 65 + f], hi, lo);
                *reinterpret_cast<uint32_t*>(sOp[p] + (((jj >> 3) * 16 + (r >> 3)) * 128 + (r & 7) * 16 + (jj & 7) * 2)) =
                    split ? lo : hi;
            }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            if (elect_one()) {
                if (MODE == 0) {
                    issue_linear_half(tmem, TL_ACCM, TL_WA, sOp[0], wdt, false);
                } else {
                    issue_linear_half(tmem, TL_ACCM, TL_WA + 32, sOp[1], wdt, false);    // W_m[:, 64:] e
                    issue_linear_half(tmem, TL_ACCM, TL_WA, sOp[0], wdt, true);          // += W_m[:, :64] agg
                }
                mma_commit(&bar);
                if (MODE == 1) issue_linear_half(tmem, TL_ACCH, TL_WB, sOp[2], wdt, false);   // W_u[:, :64] h, ahead
            }
            __syncwarp();
        }
        mbar_wait(&bar, phase); phase ^= 1u;
        tc_fence_after();
        // epilogue 1: ReLU; MODE 0 -> E rows (fp32, global); MODE 1 -> m as the next stacked operand (over the agg tile)
        float* out = eb + (size_t)outb * bs;
        for (int blk = 0; blk < (wdt >> 4); ++blk) {
            uint32_t vh[8], vl[8];
            tmem_ld_16x256b_x2(tmem_addr(tmem, 32 * warp, TL_ACCM + 16 * blk), vh);
            tmem_ld_16x256b_x2(tmem_addr(tmem, 32 * warp + 16, TL_ACCM + 16 * blk), vl);
            tmem_ld_wait();
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaxf(__uint_as_float(vh[i]) + __uint_as_float(vl[i]), 0.f);
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int n = n0 + 16 * blk + 8 * (i >> 2) + 2 * (lane & 3) + (i & 1);
                    const int ff = 16 * warp + (lane >> 2) + 8 * ((i >> 1) & 1);
                    if (n < N) out[(size_t)n * 64 + ff] = v[i];
                }
            } else {
                const int rowoff = 16 * (lane >> 2) + 4 * (lane & 3);
#pragma unroll
                for (int half = 0; half < 2; ++half)
#pragma unroll
                    for (int fr = 0; fr < 2; ++fr) {
                        uint32_t hi, lo;
                        split2(v[4 * half + 2 * fr], v[4 * half + 2 * fr + 1], hi, lo);
                        unsigned char* pp = sOp[0] + (((2 * blk + half) * 16 + 4 * warp + fr) * 128) + rowoff;
                        *reinterpret_cast<uint32_t*>(pp) = hi;
                        *reinterpret_cast<uint32_t*>(pp + 2 * 128) = lo;
                    }
            }
        }
        if (MODE == 1) {
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (warp == 0) {
                tc_fence_after();
                if (elect_one()) {
                    issue_linear_half(tmem, TL_ACCH, TL_WB + 32, sOp[0], wdt, true);     // += W_u[:, 64:] m
                    mma_commit(&bar);
                }
                __syncwarp();
            }
            mbar_wait(&bar, phase); phase ^= 1u;
            tc_fence_after();
            for (int blk = 0; blk < (wdt >> 4); ++blk) {           // epilogue 2: H' = ReLU(.) rows, fp32, global
                uint32_t vh[8], vl[8];
                tmem_ld_16x256b_x2(tmem_addr(tmem, 32 * warp, TL_ACCH + 16 * blk), vh);
                tmem_ld_16x256b_x2(tmem_addr(tmem, 32 * warp + 16, TL_ACCH + 16 * blk), vl);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int n = n0 + 16 * blk + 8 * (i >> 2) + 2 * (lane & 3) + (i & 1);
                    const int ff = 16 * warp + (lane >> 2) + 8 * ((i >> 1) & 1);
                    if (n < N) out[(size_t)n * 64 + ff] = fmaxf(__uint_as_float(vh[i]) + __uint_as_float(vl[i]), 0.f);
                }
            }
        }
        tc_fence_before();
        __syncthreads();              // staging / operand tiles and the accumulators are reused by the next item
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace

// E = ReLU(W_ef G) (layer < 0) or one message-passing layer's linears on the six-buffer episode layout of launch_mpnn_tcl
int launch_tcl_linear(const eco_graphs_t* g, const eco_mpnn_t* w, int B, float* buf, int layer, cudaStream_t st) {
    static bool attr = false;
    static int n_sm = 148;
    if (!attr) {
        ECO_CUDA(cudaFuncSetAttribute(tcl_linear_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, LSMEM));
        ECO_CUDA(cudaFuncSetAttribute(tcl_linear_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, LSMEM));
        int dev = 0;
        ECO_CUDA(cudaGetDevice(&dev));
        ECO_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        attr = true;
    }
    const long long items = (long long)B * ((g->NP + LSLAB - 1) / LSLAB);
    const int grid = (int)(items < 2LL * n_sm ? items : 2LL * n_sm);
    if (layer < 0) tcl_linear_kernel<0><<<grid, 128, LSMEM, st>>>(*g, *w, B, buf, 0);
    else tcl_linear_kernel<1><<<grid, 128, LSMEM, st>>>(*g, *w, B, buf, layer);
    ECO_LAUNCH_CHECK();
    return ECO_OK;
}

bool mpnn_tcl_supported(const eco_graphs_t* g) { return (g->reserved & 1) && g->tc_ops != nullptr; }

// OUT = scale / deg * (X1 * IMG[which1] (+ X2 * IMG[which2]));  X*, OUT: [B][*][64] fp32 with the given episode strides
int launch_tcl_contract(const eco_graphs_t* g, const int32_t* gidx, int B, const float* X1, int which1, const float* X2,
                        int which2, size_t x_stride, float* out, size_t out_stride, float scale, int edge,
                        float norm_max, cudaStream_t st) {
    static bool attr = false;
    const int smem1 = XF_BYTES + 2 * OP_BYTES, smem2 = XF_BYTES + 4 * OP_BYTES;
    if (!attr) {
        ECO_CUDA(cudaFuncSetAttribute(tcl_contract_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
        ECO_CUDA(cudaFuncSetAttribute(tcl_contract_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
        attr = true;
    }
    const int nslabs = (g->NP + SLAB - 1) / SLAB;
    const unsigned grid = (unsigned)((size_t)B * nslabs);
    if (X2)
        tcl_contract_kernel<2><<<grid, TCL_THREADS, smem2, st>>>(*g, gidx, B, X1, which1, X2, which2, x_stride, out,
                                                                 out_stride, scale, edge, norm_max);
    else
        tcl_contract_kernel<1><<<grid, TCL_THREADS, smem1, st>>>(*g, gidx, B, X1, which1, nullptr, which1, x_stride, out,
                                                                 out_stride, scale, edge, norm_max);
    ECO_LAUNCH_CHECK();
    return ECO_OK;
}

}  // namespace eco
