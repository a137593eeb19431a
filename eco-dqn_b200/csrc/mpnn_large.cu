// Kernel family 2d: MPNN forward for graphs that do not fit the resident kernel (208 < N <= 2048, couplings in {-1,0,1}).
//
// Replaces (reference, file:line)  src/networks/mpnn.py:38-74 MPNN.forward with its three layer types
//   :89-104  EdgeAndNodeEmbeddingLayer,  :114-120 UpdateNodeEmbeddingLayer,  :143-159 ReadoutLayer
// and the argmax of src/agents/dqn/dqn.py:499 (ties -> lowest index), for one observation batch.
//
// Everything between the kernels lives in HBM in ONE format ("operand tiles"), which is what the tensor cores eat:
// a plane holds a [64 features][NP vertices] activation as bf16 hi/lo pairs, 64 vertices per 16 KB tile; inside a tile
// the 128 stacked rows (r = 32q + 16s + t  <->  feature 16q + t, split s = hi / lo; same order as mpnn_tc.cu) x 64
// vertices are stored as 8x8 core matrices, core(rb, cb) at (cb*16 + rb)*128 bytes.  A tile is at once
//   * a K-major A operand  (M = stacked rows, K = vertices)       for the N x N products   X * IMG, and
//   * an MN-major B operand (K = stacked rows, N = vertices)      for the per-vertex linears W * X,
// so every kernel moves tiles with bulk copies (no conversion, no staging) and writes tiles from its epilogue.
// Per episode six planes: 0 H0, 1 H1 (ping / pong), 2 E, 3 AGG, 4 S, 5 D.
//
//   tcl_init_kernel      H0 = ReLU(W_init x);  S = R+ + R-, D = R+ - R-,  R+- = ReLU(W_x x +- w0)           CUDA cores
//   tcl_contract_kernel  AGG = scale / deg * (X1 IMG1 (+ X2 IMG2))   (edge stage: 1/2 (S |A| + D A), feature 63)  tcgen05
//   tcl_linear_kernel    E = ReLU(W_ef AGG)   |   m = ReLU(W_m [AGG ; E]),  H' = ReLU(W_u [H ; m])                 tcgen05
//   (last layer)         H' is not stored: the epilogue leaves w_r . H'_i per vertex and pooled sums per tile
//   tcl_readout_kernel   pooled readout, Q, argmax                                                            CUDA cores
#include <cuda_bf16.h>
#include <cstdlib>

#include "eco_common.cuh"
#include "tc_prims.cuh"
#include "mpnn_pack.cuh"

namespace eco {
namespace {

using namespace tc;

constexpr int TILE_V = 64;                        // vertices per operand tile
constexpr int TILE_BYTES = 128 * TILE_V * 2;      // 16 KB
constexpr int PLANES = 6;
constexpr long long TCL_PLANE_BUDGET = 4LL << 30;  // bytes of planes per pass: bounds the scratch
constexpr int PL_H0 = 0, PL_H1 = 1, PL_E = 2, PL_AGG = 3, PL_S = 4, PL_D = 5;

__host__ __device__ inline size_t plane_bytes(int NP) { return (size_t)((NP + TILE_V - 1) / TILE_V) * TILE_BYTES; }

// stacked hi row of feature f (the lo row is 16 rows = 2 core rows = 256 bytes further)
__device__ __forceinline__ int hi_row(int f) { return 32 * (f >> 4) + (f & 15); }

// ------------------------------------------------------------------------------------------------ initial embeddings
// One thread per (8 vertices, feature): mpnn.py:55 (H0) and the per-vertex half of the factorised edge stage (:89-100).
__global__ void __launch_bounds__(256)
tcl_init_kernel(const eco_graphs_t g, const eco_mpnn_t w, const float* __restrict__ xn, const float* __restrict__ xg,
                unsigned char* __restrict__ buf) {
    const int N = g.N, NP = g.NP, NB = NP >> 3;
    const int b = blockIdx.y, f = threadIdx.x & 63;
    const size_t PB = plane_bytes(NP);
    unsigned char* eb = buf + (size_t)b * PLANES * PB;
    float wi[7], we[8];
#pragma unroll
    for (int c = 0; c < 7; ++c) wi[c] = w.w_init[f * 7 + c];
#pragma unroll
    for (int c = 0; c < 8; ++c) we[c] = f < 63 ? w.w_edge[f * 8 + c] : 0.f;     // feature 63 is deg / deg_max (contraction)
    const float* x0 = xn + (size_t)b * 3 * NP;
    const float4 xgl = *reinterpret_cast<const float4*>(xg + (size_t)b * 4);
    const int r = hi_row(f);
    for (int cb = blockIdx.x * 4 + (threadIdx.x >> 6); cb < NB; cb += gridDim.x * 4) {
        float h[8], s[8], d[8], xr[3][8];
#pragma unroll
        for (int c = 0; c < 3; ++c) {                                            // NP % 16 == 0: 32-byte aligned rows
            const float4 lo = *reinterpret_cast<const float4*>(x0 + (size_t)c * NP + cb * 8);
            const float4 hi = *reinterpret_cast<const float4*>(x0 + (size_t)c * NP + cb * 8 + 4);
            xr[c][0] = lo.x; xr[c][1] = lo.y; xr[c][2] = lo.z; xr[c][3] = lo.w;
            xr[c][4] = hi.x; xr[c][5] = hi.y; xr[c][6] = hi.z; xr[c][7] = hi.w;
        }
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            const int i = cb * 8 + v;
            h[v] = s[v] = d[v] = 0.f;                                            // padding vertices stay exactly zero
            if (i < N) {
                const float X[7] = {xr[0][v], xr[1][v], xr[2][v], xgl.x, xgl.y, xgl.z, xgl.w};
                float hh = 0.f, p = 0.f;
#pragma unroll
                for (int c = 0; c < 7; ++c) {
                    hh = fmaf(wi[c], X[c], hh);
                    p = fmaf(we[1 + c], X[c], p);
                }
                const float rp = fmaxf(p + we[0], 0.f), rm = fmaxf(p - we[0], 0.f);
                h[v] = fmaxf(hh, 0.f);
                s[v] = rp + rm;
                d[v] = rp - rm;
            }
        }
        const size_t off = (size_t)(cb >> 3) * TILE_BYTES + (((cb & 7) * 16 + (r >> 3)) * 128) + (r & 7) * 16;
        auto store8 = [&](int plane, const float (&v)[8]) {
            uint4 hi, lo;
            split2(v[0], v[1], hi.x, lo.x);
            split2(v[2], v[3], hi.y, lo.y);
            split2(v[4], v[5], hi.z, lo.z);
            split2(v[6], v[7], hi.w, lo.w);
            unsigned char* p = eb + (size_t)plane * PB + off;
            *reinterpret_cast<uint4*>(p) = hi;
            *reinterpret_cast<uint4*>(p + 256) = lo;
        };
        store8(PL_H0, h);
        store8(PL_S, s);
        store8(PL_D, d);
    }
}

// Epilogue store of the values of two adjacent vertices (even column first) of one feature row into a tile plane:
// the lane holds feature 16q + lane/4 + 8 fr and vertices nbase + 2 (lane & 3) + {0, 1}, nbase a multiple of 8.
__device__ __forceinline__ void store_pair(unsigned char* plane, int nbase, int q, int fr, int lane, float va, float vb) {
    uint32_t hi, lo;
    split2(va, vb, hi, lo);
    unsigned char* p = plane + (size_t)(nbase >> 6) * TILE_BYTES + ((((nbase & 63) >> 3) * 16 + 4 * q + fr) * 128) +
                       16 * (lane >> 2) + 4 * (lane & 3);
    *reinterpret_cast<uint32_t*>(p) = hi;
    *reinterpret_cast<uint32_t*>(p + 256) = lo;
}

// ------------------------------------------------------------------------------------------------ N x N products
// Persistent CTAs (one per SM) over (episode, 256-column slab) items; K is walked in 64-vertex panels, PAIRS operand
// pairs one after the other.  Warp 0 (one lane) feeds a four-stage ring with bulk copies (activation tile + the matching
// panel of the graph's bf16 operand image, graph_prepare.cu: tc_ops), warp 1 (one lane) issues tcgen05.mma into one of
// two 256-column TMEM accumulators, warps 2-5 (one per TMEM lane quadrant) drain the other one: the epilogue of a slab
// overlaps the main loop of the next.
constexpr int CW = 256;
constexpr int CSTAGES = 4;
constexpr int CB_BYTES = TILE_V * CW * 2;                 // 32 KB: 64 k x 256 columns
constexpr int CSTAGE_BYTES = TILE_BYTES + CB_BYTES;
constexpr int CSMEM = CSTAGES * CSTAGE_BYTES;
constexpr int CTHREADS = 192;

template <int PAIRS>
__global__ void __launch_bounds__(CTHREADS, 1)
tcl_contract_kernel(const eco_graphs_t g, const int32_t* __restrict__ graph_idx, const int B, unsigned char* __restrict__ buf,
                    const int src1, const int which1, const int src2, const int which2, const int dst,
                    const float scale, const int edge, const float norm_max) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[CSTAGES], empty[CSTAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) float s_rdeg[2][CW], s_dn[2][CW];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int N = g.N, NP = g.NP, NB = NP >> 3;
    const int nslabs = (NP + CW - 1) / CW, nitems = B * nslabs;
    const size_t PB = plane_bytes(NP);
    const int npanels = (NP + TILE_V - 1) / TILE_V, units = PAIRS * npanels;

    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        for (int s = 0; s < CSTAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 128); }
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (warp == 0) {
        if (elect_one()) {
            int uc = 0;                                    // units fed so far (ring position)
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const int b = item / nslabs, n0 = (item % nslabs) * CW, w = min(CW, NP - n0);   // NP % 16 == 0, so is w
                const int gi = graph_idx[b], run = (w >> 3) * 128;    // run: bytes of one 8-vertex K group of the panel
                const unsigned char* eb = buf + (size_t)b * PLANES * PB;
                for (int u = 0; u < units; ++u, ++uc) {
                    const int s = uc % CSTAGES, p = u / npanels, kp = u % npanels;
                    const int k0 = kp * TILE_V, kg = min(TILE_V, NP - k0) >> 3;
                    if (uc >= CSTAGES) mbar_wait(&empty[s], (uint32_t)((uc / CSTAGES - 1) & 1));
                    unsigned char* sa = smem + s * CSTAGE_BYTES;
                    unsigned char* sb = sa + TILE_BYTES;
                    const uint16_t* img = g.tc_ops + ((size_t)gi * 2 + (p ? which2 : which1)) * NP * NP;
                    mbar_expect_tx(&full[s], (uint32_t)(kg * (2048 + run)));
                    bulk_g2s(sa, eb + (size_t)(p ? src2 : src1) * PB + (size_t)kp * TILE_BYTES, kg * 2048, &full[s]);
                    // the panel's K groups of this slab are contiguous in the image (eco_common.cuh: tc_image_core)
                    bulk_g2s(sb, reinterpret_cast<const unsigned char*>(img) + tc_image_core(NB, k0 >> 3, n0 >> 3), kg * run, &full[s]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            int uc = 0, k = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++k) {
                const int w = min(CW, NP - (item % nslabs) * CW), run = (w >> 3) * 128;
                const uint32_t idesc = instr_desc_bf16(128, w, false, false);
                const uint32_t acc = tmem + (uint32_t)(k & 1) * CW;
                if (k >= 2) { mbar_wait(&acc_empty[k & 1], (uint32_t)(((k >> 1) - 1) & 1)); tc_fence_after(); }
                for (int u = 0; u < units; ++u, ++uc) {
                    const int s = uc % CSTAGES, kp = u % npanels;
                    const int kw = min(TILE_V, NP - kp * TILE_V);
                    mbar_wait(&full[s], (uint32_t)((uc / CSTAGES) & 1));
                    tc_fence_after();
                    const uint64_t ad = smem_desc(smem_u32(smem + s * CSTAGE_BYTES), 2048, 128);
                    const uint64_t bd = smem_desc(smem_u32(smem + s * CSTAGE_BYTES + TILE_BYTES), run, 128);
                    for (int ks = 0; ks < (kw >> 4); ++ks)
                        mma_ss(acc, ad + (uint64_t)ks * (4096 >> 4), bd + (uint64_t)ks * ((2 * run) >> 4), idesc, u > 0 || ks > 0);
                    mma_commit(&empty[s]);
                }
                mma_commit(&acc_full[k & 1]);
            }
        }
        __syncwarp();
    } else {
        // epilogue warps: (hi + lo rows) * scale / deg; feature 63 of the edge stage = deg / deg_max (mpnn.py:102)
        const int q = warp & 3, et = tid - 64;             // TMEM lane quadrant of this warp; index among the 128 threads
        const bool f63 = edge && q == 3 && (lane >> 2) == 7;        // this lane's fr = 1 row is feature 63
        int k = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++k) {
            const int b = item / nslabs, n0 = (item % nslabs) * CW, w = min(CW, NP - n0);
            const int gi = graph_idx[b], a = k & 1;
            const float dmax = norm_max > 0.f ? norm_max : (norm_max < 0.f ? (float)max(g.gstat[(size_t)gi * 4], 1) : *g.dmax);
            for (int i = et; i < w; i += 128) {            // 0 for padding vertices
                const bool ok = n0 + i < N;
                const float d = ok ? g.deg[(size_t)gi * NP + n0 + i] : 1.f;
                s_rdeg[a][i] = ok ? scale / d : 0.f;
                s_dn[a][i] = ok ? d / dmax : 0.f;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            mbar_wait(&acc_full[a], (uint32_t)((k >> 1) & 1));
            tc_fence_after();
            unsigned char* ob = buf + ((size_t)b * PLANES + dst) * PB;
#pragma unroll 2
            for (int blk = 0; blk < (w >> 4); ++blk) {
                uint32_t vh[8], vl[8];
                tmem_ld_16x256b_x2(tmem_addr(tmem, 32 * q, a * CW + 16 * blk), vh);
                tmem_ld_16x256b_x2(tmem_addr(tmem, 32 * q + 16, a * CW + 16 * blk), vl);
                tmem_ld_wait();
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int c = 16 * blk + 8 * half + 2 * (lane & 3);
                    const float2 rd = *reinterpret_cast<const float2*>(&s_rdeg[a][c]);
                    const float2 dn = *reinterpret_cast<const float2*>(&s_dn[a][c]);
#pragma unroll
                    for (int fr = 0; fr < 2; ++fr) {
                        const int i = 4 * half + 2 * fr;
                        float va = (__uint_as_float(vh[i]) + __uint_as_float(vl[i])) * rd.x;
                        float vb = (__uint_as_float(vh[i + 1]) + __uint_as_float(vl[i + 1])) * rd.y;
                        if (fr == 1 && f63) { va = dn.x; vb = dn.y; }
                        store_pair(ob, n0 + 16 * blk + 8 * half, q, fr, lane, va, vb);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[a]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ per-vertex linears
// Persistent CTAs (two per SM) over (episode, 64-vertex tile) items: the weights sit in TMEM as A operands for the whole
// launch (packed bf16 hi/lo rows, eco_mpnn_pack), the item's tiles arrive by bulk copies one item ahead, the message m
// goes back to shared memory as the next B operand (over the AGG tile), H' / E leave as tiles.
constexpr uint32_t TL_WA = 0, TL_WB = 64, TL_ACCM = 128, TL_ACCH = 192;
constexpr int LSMEM = 2 * 3 * TILE_BYTES;

// 8 MMAs: acc (+)= W[:, 64-feature group at TMEM column tw] * X, X a tile used as MN-major B operand
__device__ __forceinline__ void issue_linear_half(uint32_t tmem, uint32_t acc_col, uint32_t tw, const unsigned char* x, int width,
                                                  bool accumulate) {
    // (the lo rows of X meet W_hi only: M = 64 covers TMEM lanes 32q + 0..15, the hi rows of the stacked order -- mpnn_tc.cu)
    const uint32_t idesc = instr_desc_bf16(128, width, false, true);
    const uint32_t idesc_lo = instr_desc_bf16(64, width, false, true);
    const uint64_t d = smem_desc(smem_u32(x), /*LBO (k groups)*/ 128, /*SBO (vertex groups)*/ 2048);
#pragma unroll
    for (int i = 0; i < 8; ++i)     // i = 2*kq + s: features 16kq..16kq+15, split s
        mma_ts(tmem + acc_col, tmem + tw + 8 * (i >> 1), d + (uint64_t)(16 * i), (i & 1) ? idesc_lo : idesc, accumulate || i > 0);
}

template <int MODE>
__global__ void __launch_bounds__(128, 2)
tcl_linear_kernel(const eco_graphs_t g, const eco_mpnn_t w, const int B, unsigned char* __restrict__ buf, const int layer,
                  float* __restrict__ qpart, float* __restrict__ ppart) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[2], bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) float sq[4][TILE_V], sp[64];              // MODE 2: per-warp Q partial sums, per-feature pooled sums of the tile
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int NP = g.NP;
    const size_t PB = plane_bytes(NP);
    const int ntiles = (NP + TILE_V - 1) / TILE_V, nitems = B * ntiles;
    const uint32_t* pk = reinterpret_cast<const uint32_t*>(w.packed);
    constexpr int NIN = MODE == 0 ? 1 : 3;
    const int in_plane[3] = {PL_AGG, PL_E, (layer & 1) ? PL_H1 : PL_H0};
    const int out_plane = MODE == 0 ? PL_E : ((layer & 1) ? PL_H0 : PL_H1);

    if (warp == 0) tmem_alloc(&tmem_base_s, 256);
    if (tid == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_init(&bar, 1); fence_mbar_init(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    auto fetch = [&](int item, int s) {                   // one thread: the item's tiles -> stage s
        const int b = item / ntiles, t = item % ntiles;
        const uint32_t bytes = (uint32_t)(min(TILE_V, NP - t * TILE_V) >> 3) * 2048;
        mbar_expect_tx(&full[s], NIN * bytes);
        for (int p = 0; p < NIN; ++p)
            bulk_g2s(smem + (s * 3 + p) * TILE_BYTES, buf + ((size_t)b * PLANES + in_plane[p]) * PB + (size_t)t * TILE_BYTES,
                     bytes, &full[s]);
    };
    if (tid == 0 && (int)blockIdx.x < nitems) fetch(blockIdx.x, 0);
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
    // MODE 2: readout weights of this lane's two feature rows (16 warp + lane/4, + 8)
    const float wr0 = MODE == 2 ? w.w_read[64 + 16 * warp + (lane >> 2)] : 0.f;
    const float wr1 = MODE == 2 ? w.w_read[64 + 16 * warp + (lane >> 2) + 8] : 0.f;
    uint32_t phase = 0;
    int it = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
        const int s = it & 1;
        const int b = item / ntiles, n0 = (item % ntiles) * TILE_V, wdt = min(TILE_V, NP - n0);
        unsigned char* sAgg = smem + (s * 3 + 0) * TILE_BYTES;
        unsigned char* sE = smem + (s * 3 + 1) * TILE_BYTES;
        unsigned char* sH = smem + (s * 3 + 2) * TILE_BYTES;
        if (warp == 0) {
            if (elect_one()) {
                if (item + (int)gridDim.x < nitems) fetch(item + gridDim.x, s ^ 1);   // the other stage is idle since the last item
                mbar_wait(&full[s], (uint32_t)((it >> 1) & 1));
                tc_fence_after();
                if (MODE == 0) {
                    issue_linear_half(tmem, TL_ACCM, TL_WA, sAgg, wdt, false);
                } else {
                    issue_linear_half(tmem, TL_ACCM, TL_WA + 32, sE, wdt, false);       // W_m[:, 64:] e
                    issue_linear_half(tmem, TL_ACCM, TL_WA, sAgg, wdt, true);           // += W_m[:, :64] agg
                }
                mma_commit(&bar);
                if (MODE >= 1) issue_linear_half(tmem, TL_ACCH, TL_WB, sH, wdt, false);  // W_u[:, :64] h, ahead
            }
            __syncwarp();
        }
        mbar_wait(&bar, phase); phase ^= 1u;
        tc_fence_after();
        unsigned char* ob = buf + ((size_t)b * PLANES + out_plane) * PB;
        // epilogue 1: ReLU; MODE 0 -> E tile (global); MODE 1 -> m as the next B operand (over the AGG tile)
        for (int blk = 0; blk < (wdt >> 4); ++blk) {
            uint32_t vh[8], vl[8];
            tmem_ld_16x256b_x2(tmem_addr(tmem, 32 * warp, TL_ACCM + 16 * blk), vh);
            tmem_ld_16x256b_x2(tmem_addr(tmem, 32 * warp + 16, TL_ACCM + 16 * blk), vl);
            tmem_ld_wait();
#pragma unroll
            for (int half = 0; half < 2; ++half)
#pragma unroll
                for (int fr = 0; fr < 2; ++fr) {
                    const int i = 4 * half + 2 * fr;
                    const float va = fmaxf(__uint_as_float(vh[i]) + __uint_as_float(vl[i]), 0.f);
                    const float vb = fmaxf(__uint_as_float(vh[i + 1]) + __uint_as_float(vl[i + 1]), 0.f);
                    if (MODE == 0) store_pair(ob, n0 + 16 * blk + 8 * half, warp, fr, lane, va, vb);
                    else store_pair(sAgg, 16 * blk + 8 * half, warp, fr, lane, va, vb);
                }
        }
        if (MODE >= 1) {
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (warp == 0) {
                tc_fence_after();
                if (elect_one()) {
                    issue_linear_half(tmem, TL_ACCH, TL_WB + 32, sAgg, wdt, true);       // += W_u[:, 64:] m
                    mma_commit(&bar);
                }
                __syncwarp();
            }
            mbar_wait(&bar, phase); phase ^= 1u;
            tc_fence_after();
            float ps0 = 0.f, ps1 = 0.f;
            for (int blk = 0; blk < (wdt >> 4); ++blk) {           // epilogue 2: H' = ReLU(.) tile
                uint32_t vh[8], vl[8];
                tmem_ld_16x256b_x2(tmem_addr(tmem, 32 * warp, TL_ACCH + 16 * blk), vh);
                tmem_ld_16x256b_x2(tmem_addr(tmem, 32 * warp + 16, TL_ACCH + 16 * blk), vl);
                tmem_ld_wait();
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fmaxf(__uint_as_float(vh[i]) + __uint_as_float(vl[i]), 0.f);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    if (MODE == 1) {
                        store_pair(ob, n0 + 16 * blk + 8 * half, warp, 0, lane, v[4 * half], v[4 * half + 1]);
                        store_pair(ob, n0 + 16 * blk + 8 * half, warp, 1, lane, v[4 * half + 2], v[4 * half + 3]);
                    } else {
                        // last layer: H' only feeds the readout (mpnn.py:143-159); keep its per-vertex and pooled sums
                        ps0 += v[4 * half] + v[4 * half + 1];
                        ps1 += v[4 * half + 2] + v[4 * half + 3];
                        float qa = fmaf(wr0, v[4 * half], wr1 * v[4 * half + 2]);
                        float qb = fmaf(wr0, v[4 * half + 1], wr1 * v[4 * half + 3]);
#pragma unroll
                        for (int o = 4; o < 32; o <<= 1) {
                            qa += __shfl_xor_sync(0xffffffffu, qa, o);
                            qb += __shfl_xor_sync(0xffffffffu, qb, o);
                        }
                        if ((lane >> 2) == 0)
                            *reinterpret_cast<float2*>(&sq[warp][16 * blk + 8 * half + 2 * (lane & 3)]) = make_float2(qa, qb);
                    }
                }
            }
            if (MODE == 2) {
#pragma unroll
                for (int o = 1; o < 4; o <<= 1) {
                    ps0 += __shfl_xor_sync(0xffffffffu, ps0, o);
                    ps1 += __shfl_xor_sync(0xffffffffu, ps1, o);
                }
                if ((lane & 3) == 0) { sp[16 * warp + (lane >> 2)] = ps0; sp[16 * warp + (lane >> 2) + 8] = ps1; }
                __syncthreads();
                if (tid < wdt) qpart[(size_t)b * NP + n0 + tid] = (sq[0][tid] + sq[1][tid]) + (sq[2][tid] + sq[3][tid]);
                if (tid < 64) ppart[((size_t)b * ntiles + n0 / TILE_V) * 64 + tid] = sp[tid];
            }
        }
        tc_fence_before();
        __syncthreads();              // this stage and the accumulators are reused
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------------ readout
// One CTA per episode: mpnn.py:143-159 from the sums the last layer left (per-vertex w_r[64:] . H'_i, per-tile pooled
// sums), then argmax (ties -> lowest index).
constexpr int RD_THREADS = 256;

__global__ void __launch_bounds__(RD_THREADS)
tcl_readout_kernel(const eco_graphs_t g, const eco_mpnn_t w, const float* __restrict__ qpart, const float* __restrict__ ppart,
                   float* __restrict__ q_out, int32_t* __restrict__ act_out) {
    __shared__ float pooled[64];
    __shared__ float c0_s;
    __shared__ float red_val[RD_THREADS / 32];
    __shared__ int red_idx[RD_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = g.N, NP = g.NP, b = blockIdx.x;
    const int ntiles = (NP + TILE_V - 1) / TILE_V;
    if (tid < 64) {
        float s = 0.f;
        for (int t = 0; t < ntiles; ++t) s += ppart[((size_t)b * ntiles + t) * 64 + tid];
        pooled[tid] = s / (float)N;
    }
    __syncthreads();
    if (warp == 0) {
        float c = 0.f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int f = lane + 32 * half;
            float p = 0.f;
            for (int k = 0; k < 64; ++k) p = fmaf(w.w_pool[f * 64 + k], pooled[k], p);
            c = fmaf(w.w_read[f], fmaxf(p, 0.f), c);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) c0_s = c + w.b_read[0];
    }
    __syncthreads();
    const float c0 = c0_s;
    float best_v = -INFINITY;
    int best_i = 0x7fffffff;
    for (int i = tid; i < N; i += RD_THREADS) {
        const float v = qpart[(size_t)b * NP + i] + c0;
        if (q_out) q_out[(size_t)b * NP + i] = v;
        if (v > best_v) { best_v = v; best_i = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best_v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ov > best_v || (ov == best_v && oi < best_i)) { best_v = ov; best_i = oi; }
    }
    if (lane == 0) { red_val[warp] = best_v; red_idx[warp] = best_i; }
    __syncthreads();
    if (tid == 0 && act_out) {
        float bv = red_val[0];
        int bi = red_idx[0];
        for (int ww = 1; ww < RD_THREADS / 32; ++ww)
            if (red_val[ww] > bv || (red_val[ww] == bv && red_idx[ww] < bi)) { bv = red_val[ww]; bi = red_idx[ww]; }
        act_out[b] = bi;
    }
}

}  // namespace

// Episodes per pass of the ten launches.  One pass unless the planes would exceed TCL_PLANE_BUDGET: keeping the planes
// L2-resident with small passes was measured and loses (ER-500, B = 1024: one pass 0.81 ms, 512 episodes per pass 0.90,
// 128 per pass 1.34 -- ten launch tails per pass cost more than the L2 hits save).  ECO_TCL_CHUNK overrides.
static int tcl_chunk(int B, int NP) {
    static const int forced = getenv("ECO_TCL_CHUNK") ? atoi(getenv("ECO_TCL_CHUNK")) : 0;
    long long c = forced > 0 ? forced : (long long)TCL_PLANE_BUDGET / (long long)(PLANES * plane_bytes(NP));
    if (c < 1) c = 1;
    return c < B ? (int)c : B;
}

// scratch: `chunk` episodes x 6 planes of operand tiles, then the last layer's readout sums [chunk][NP] + [chunk][tiles][64]
size_t mpnn_tcl_scratch_bytes(int B, int N) {
    const int NP = padded_n(N);
    const size_t c = (size_t)tcl_chunk(B, NP);
    return align256(c * PLANES * plane_bytes(NP) + sizeof(float) * (c * NP + c * ((NP + TILE_V - 1) / TILE_V) * 64));
}

int launch_mpnn_tcl(const eco_graphs_t* g, const eco_mpnn_t* w, int B, const int32_t* gidx, const float* xn,
                    const float* xg, float norm_max, float* q, int32_t* actions, void* scratch, cudaStream_t st) {
    static unsigned long long attr = 0;
    const int n_sm = device_sm_count();
    if (first_use_on_device(&attr)) {
        ECO_CUDA(cudaFuncSetAttribute(tcl_contract_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, CSMEM));
        ECO_CUDA(cudaFuncSetAttribute(tcl_contract_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, CSMEM));
        ECO_CUDA(cudaFuncSetAttribute(tcl_linear_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, LSMEM));
        ECO_CUDA(cudaFuncSetAttribute(tcl_linear_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, LSMEM));
        ECO_CUDA(cudaFuncSetAttribute(tcl_linear_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, LSMEM));
    }
    unsigned char* buf = (unsigned char*)scratch;
    const int NP = g->NP, NB = NP >> 3;
    const int ntiles = (NP + TILE_V - 1) / TILE_V, nslabs = (NP + CW - 1) / CW;
    const int chunk = tcl_chunk(B, NP);
    float* qpart = reinterpret_cast<float*>(buf + (size_t)chunk * PLANES * plane_bytes(NP));
    float* ppart = qpart + (size_t)chunk * NP;
    prof_begin(ECO_PROF_MPNN, st);
    // Episodes go through the ten launches `chunk` at a time (normally all at once); the planes are reused by every pass.
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int nb = B - b0 < chunk ? B - b0 : chunk;
        const int32_t* gi = gidx + b0;
        const long long items = (long long)nb * ntiles;
        const int lgrid = (int)(items < 2LL * n_sm ? items : 2LL * n_sm);
        const long long citems = (long long)nb * nslabs;
        const int cgrid = (int)(citems < n_sm ? citems : n_sm);
        tcl_init_kernel<<<dim3((NB + 3) / 4, nb), 256, 0, st>>>(*g, *w, xn + (size_t)b0 * 3 * NP, xg + (size_t)b0 * 4, buf);
        ECO_LAUNCH_CHECK();
        // g = (S |A| + D A) / (2 deg), feature 63 = deg / deg_max
        tcl_contract_kernel<2><<<cgrid, CTHREADS, CSMEM, st>>>(*g, gi, nb, buf, PL_S, 1, PL_D, 0, PL_AGG, 0.5f, 1, norm_max);
        ECO_LAUNCH_CHECK();
        tcl_linear_kernel<0><<<lgrid, 128, LSMEM, st>>>(*g, *w, nb, buf, 0, nullptr, nullptr);
        ECO_LAUNCH_CHECK();
        for (int l = 0; l < 3; ++l) {
            tcl_contract_kernel<1><<<cgrid, CTHREADS, CSMEM, st>>>(*g, gi, nb, buf, (l & 1) ? PL_H1 : PL_H0, 0, 0, 0, PL_AGG, 1.f, 0, norm_max);
            ECO_LAUNCH_CHECK();
            if (l < 2) tcl_linear_kernel<1><<<lgrid, 128, LSMEM, st>>>(*g, *w, nb, buf, l, nullptr, nullptr);
            else tcl_linear_kernel<2><<<lgrid, 128, LSMEM, st>>>(*g, *w, nb, buf, l, qpart, ppart);   // H' only as readout sums
            ECO_LAUNCH_CHECK();
        }
        tcl_readout_kernel<<<nb, RD_THREADS, 0, st>>>(*g, *w, qpart, ppart, q ? q + (size_t)b0 * NP : nullptr,
                                                      actions ? actions + b0 : nullptr);
        ECO_LAUNCH_CHECK();
    }
    prof_end(ECO_PROF_MPNN, st);
    return ECO_OK;
}

}  // namespace eco
