// Graph-set preparation: padding and the per-graph scorer constants.
//
// Replaces (reference, file:line)
//   src/envs/score_solver.py:353-357  set_quality_normalizer  qn  = max(1, sum_{J>0} J / 2)
//   src/envs/score_solver.py:359-365  set_lower_bound         lb  = min(0, sum_{J<0} J / 2)
//   src/envs/score_solver.py:367-375  set_max_local_reward    mlr = max over NON-ZERO weighted degrees
//   src/networks/mpnn.py:34-38        get_normalisation       deg_i = max(1, #{j : J_ij != 0})
// All sums are integer (int8 couplings), hence exact and identical to the reference's fp64 sums.
#include <cuda_bf16.h>

#include "eco_common.cuh"

namespace eco {

__global__ void graph_pad_kernel(const int8_t* __restrict__ dense, int8_t* __restrict__ J, int G, int N, int NP) {
    // `dense` holds G graphs; J points at the first padded slot they go to
    // one thread per 16-byte chunk of the padded layout
    const size_t chunks_per_row = NP / 16;
    const size_t total = (size_t)G * NP * chunks_per_row;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int c = idx % chunks_per_row;
        const size_t row = idx / chunks_per_row;
        const int i = row % NP;
        const size_t gi = row / NP;
        union { int8_t b[16]; uint4 v; } u;
        u.v = make_uint4(0, 0, 0, 0);
        if (i < N) {
            const int8_t* src = dense + (gi * N + i) * (size_t)N + c * 16;
#pragma unroll
            for (int k = 0; k < 16; ++k)
                if (c * 16 + k < N) u.b[k] = src[k];
        }
        *reinterpret_cast<uint4*>(J + row * NP + c * 16) = u.v;
    }
}

// one CTA per graph, one warp per row (strided)
__global__ void __launch_bounds__(256) graph_prepare_kernel(eco_graphs_t g, int first) {
    const int gi = first + blockIdx.x;
    const int N = g.N, NP = g.NP;
    const int8_t* J = g.J + (size_t)gi * NP * NP;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;

    __shared__ int s_sum[8], s_pos[8], s_neg[8], s_mlr[8], s_maxdeg[8], s_nnz[8], s_maxabs[8], s_flags[8];
    int t_sum = 0, t_pos = 0, t_neg = 0, t_mlr = INT_MIN, t_maxdeg = 0, t_nnz = 0, t_maxabs = 0, t_flags = 0;
    const bool min_cut = (g.reserved & ECO_GRAPHS_MIN_CUT) != 0;     // score_solver.py:423-505

    for (int i = warp; i < NP; i += nwarp) {
        int rs = 0, ra = 0, rp = 0, rn = 0, cnt = 0, bad = 0;
        for (int j = lane; j < NP; j += 32) {
            const int v = J[(size_t)i * NP + j];
            rs += v;
            ra += abs(v);
            rp += v > 0 ? v : 0;
            rn += v < 0 ? v : 0;
            cnt += v != 0;
            bad |= (v > 1 || v < -1) ? 1 : 0;
            if (i < N && j < N) bad |= (J[(size_t)j * NP + i] != v) ? 4 : 0;
            if (i == j && v != 0) bad |= 4;
            if ((i >= N || j >= N) && v != 0) bad |= 4;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            rs += __shfl_xor_sync(0xffffffffu, rs, o);
            ra += __shfl_xor_sync(0xffffffffu, ra, o);
            rp += __shfl_xor_sync(0xffffffffu, rp, o);
            rn += __shfl_xor_sync(0xffffffffu, rn, o);
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            bad |= __shfl_xor_sync(0xffffffffu, bad, o);
        }
        if (lane == 0) g.deg[(size_t)gi * NP + i] = (float)max(cnt, 1);
        t_sum += rs;
        t_pos += rp;
        t_neg += rn;
        t_nnz += cnt;
        t_flags |= bad;
        t_maxdeg = max(t_maxdeg, cnt);
        t_maxabs = max(t_maxabs, ra);
        if (i < N && rs != 0) t_mlr = max(t_mlr, min_cut ? -rs : rs);   // max non-zero score-mask entry of the all -1 state
    }
    if (lane == 0) {
        s_sum[warp] = t_sum; s_pos[warp] = t_pos; s_neg[warp] = t_neg; s_mlr[warp] = t_mlr;
        s_maxdeg[warp] = t_maxdeg; s_nnz[warp] = t_nnz; s_maxabs[warp] = t_maxabs; s_flags[warp] = t_flags;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int sum = 0, pos = 0, neg = 0, mlr = INT_MIN, maxdeg = 0, nnz = 0, maxabs = 0, flags = 0;
        for (int w = 0; w < nwarp; ++w) {
            sum += s_sum[w]; pos += s_pos[w]; neg += s_neg[w]; mlr = max(mlr, s_mlr[w]);
            maxdeg = max(maxdeg, s_maxdeg[w]); nnz += s_nnz[w]; maxabs = max(maxabs, s_maxabs[w]);
            flags |= s_flags[w];
        }
        if (mlr == INT_MIN) { flags |= 2; mlr = 1; }
        double* gs = g.gscal + (size_t)gi * 4;
        gs[0] = (double)mlr;
        gs[1] = min_cut ? fmax(1.0, fabs((double)neg)) : fmax(1.0, (double)pos / 2.0);   // :439-443 / :353-357
        gs[2] = fmin(0.0, (double)neg / 2.0);
        gs[3] = (double)sum;
        int32_t* st = g.gstat + (size_t)gi * 4;
        st[0] = maxdeg; st[1] = nnz; st[2] = maxabs; st[3] = flags;
        s_sum[0] = mlr;
        s_pos[0] = pos;
        s_neg[0] = neg;
    }
    __syncthreads();
    // observable row 1 as a table: immediate_quality_changes / max_local_reward in fp64, then the driver's fp32 cast
    // (spinsystem.py:490, experiments/utils.py:174), for every cut change k = -NP..NP a +-1 graph can produce
    const double mlr_d = (double)s_sum[0];
    float* tab = g.gain_tab + (size_t)gi * tab_stride(NP);
    const double qn_d = min_cut ? fmax(1.0, fabs((double)s_neg[0])) : fmax(1.0, (double)s_pos[0] / 2.0);
    double* dtab = g.dn_tab + (size_t)gi * tab_stride(NP);
    for (int k = threadIdx.x; k <= 2 * NP; k += blockDim.x) {
        tab[k] = (float)__ddiv_rn((double)(k - NP), mlr_d);
        dtab[k] = __ddiv_rn((double)(k - NP), qn_d);                  // delta_score / quality normaliser (spinsystem.py:394)
    }
    // bf16 operand images of J and |J| for the tensor-core MPNN (mpnn_tc.cu): element (r, c) of the K-major operand
    // lives in core matrix (rb = r / 8, cb = c / 8) at byte tc_image_core(NP/8, cb, rb) + (r % 8) * 16 + (c % 8) * 2; for
    // NP <= 256 that is ((cb * NP/8 + rb) * 8 + r % 8) * 16 + ..., exactly the resident kernel's shared-memory layout, so
    // an episode fetches each image with one bulk copy.  int8 is exact in bf16.
    if (g.tc_ops != nullptr) {
        const int NB = NP >> 3;
        uint4* img_a = reinterpret_cast<uint4*>(g.tc_ops + (size_t)gi * 2 * NP * NP);
        uint4* img_abs = img_a + (size_t)NP * NP / 8;
        for (int idx = threadIdx.x; idx < NB * NB * 8; idx += blockDim.x) {
            const int r7 = idx & 7, rb = (idx >> 3) % NB, cb = (idx >> 3) / NB;
            const int8_t* src = J + (size_t)(rb * 8 + r7) * NP + cb * 8;
            const uint2 raw = *reinterpret_cast<const uint2*>(src);
            uint32_t wa[4], wb[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t word = e < 2 ? raw.x : raw.y;
                const int v0 = (int)(int8_t)(word >> (16 * (e & 1))), v1 = (int)(int8_t)(word >> (16 * (e & 1) + 8));
                wa[e] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn((float)v0)) |
                        ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn((float)v1)) << 16);
                wb[e] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn((float)abs(v0))) |
                        ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn((float)abs(v1))) << 16);
            }
            const size_t dst = tc_image_core(NB, cb, rb) / 16 + r7;
            img_a[dst] = make_uint4(wa[0], wa[1], wa[2], wa[3]);
            img_abs[dst] = make_uint4(wb[0], wb[1], wb[2], wb[3]);
        }
    }
}

// Edge lists -> the padded dense int8 layout (eco_graphs_load_edges_dev).  One thread per entry; the graph of an entry is
// found by bisection over the (count + 1) offsets.  Out-of-range vertices raise *err instead of writing.
__global__ void graph_scatter_edges_kernel(int8_t* __restrict__ J, const int N, const int NP, const int count,
                                           const int64_t* __restrict__ offsets, const int32_t* __restrict__ rows,
                                           const int32_t* __restrict__ cols, const int8_t* __restrict__ wts,
                                           const long long n_entries, const int symmetric, int* __restrict__ err) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_entries; e += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = count;                       // largest k with offsets[k] <= e
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (offsets[mid] <= e) lo = mid; else hi = mid;
        }
        const int i = rows[e], j = cols[e];
        if (i < 0 || i >= N || j < 0 || j >= N) { atomicExch(err, 1); continue; }
        int8_t* Jk = J + (size_t)lo * NP * NP;
        Jk[(size_t)i * NP + j] = wts[e];
        if (symmetric) Jk[(size_t)j * NP + i] = wts[e];
    }
}

int launch_graph_scatter_edges(const eco_graphs_t* g, int first, int count, const int64_t* offsets, const int32_t* rows,
                               const int32_t* cols, const int8_t* wts, long long n_entries, int symmetric, int* err_dev,
                               cudaStream_t st) {
    if (n_entries == 0) return ECO_OK;
    const long long want = (n_entries + 255) / 256;
    const int blocks = (int)(want < 148 * 16 ? want : 148 * 16);
    graph_scatter_edges_kernel<<<blocks, 256, 0, st>>>(g->J + (size_t)first * g->NP * g->NP, g->N, g->NP, count, offsets, rows,
                                                       cols, wts, n_entries, symmetric, err_dev);
    ECO_LAUNCH_CHECK();
    return ECO_OK;
}

int launch_graph_pad(const eco_graphs_t* g, const int8_t* dense_dev, int first, int count, cudaStream_t st) {
    const size_t total = (size_t)count * g->NP * (g->NP / 16);
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    graph_pad_kernel<<<blocks, 256, 0, st>>>(dense_dev, g->J + (size_t)first * g->NP * g->NP, count, g->N, g->NP);
    ECO_LAUNCH_CHECK();
    return ECO_OK;
}

int launch_graph_prepare(const eco_graphs_t* g, int first, int count, cudaStream_t st) {
    graph_prepare_kernel<<<count, 256, 0, st>>>(*g, first);
    ECO_LAUNCH_CHECK();
    return ECO_OK;
}

}  // namespace eco
