// The optimizer step of the DQN update with its state on the device, and -- for data-parallel training on the GPUs of one
// NVSwitch box -- fused with the gradient all-reduce over peer memory.
//
// Replaces (reference, file:line) src/agents/dqn/dqn.py:449 `self.optimizer.step()` (torch.optim.Adam, :212); the reference
// has no multi-GPU path, the data-parallel form is SURVEY.md section 8(e): every rank owns its environments and replay
// shard, gradients are averaged once per update, every rank applies the same Adam step.
//
//   eco_mpnn_adam_dev   Adam with the step counter and the learning rate read from DEVICE memory (the counter is advanced
//                       by the kernel), so the whole update -- TD target, gradient kernels, Adam, operand re-pack -- can be
//                       captured in one CUDA graph and replayed without a host round trip.
//   eco_dp_*            one kernel per update that (a) publishes this rank's gradient in a buffer its peers can read, (b)
//                       waits until every peer has done the same (flags written over NVLink with system-scope release /
//                       acquire), (c) sums the W gradients straight out of peer memory in rank order -- so all ranks get
//                       bit-identical sums -- and (d) applies Adam to its own copy of the parameters.  No NCCL call, no
//                       second pass over the gradient: 234 kB per peer cross NVLink while the update is computed.
//
// Exchange buffers are cudaMalloc'ed here and shared with cudaIpc handles (one process per GPU).  Two gradient slots
// alternate between updates: a rank can only reach update e after every rank finished update e - 1, i.e. nobody still
// reads slot (e & 1) from update e - 2, so no "done" handshake is needed.
#include <cuda_runtime.h>
#include <math.h>
#include <new>

#include "eco_common.cuh"

struct eco_dp {
    int world, rank;
    int n_sm;
    unsigned char* local;                 // [2][N_PARAMS] float slots | flags[ECO_DP_MAX_RANKS] int | arrive counter (u64)
    unsigned char* peer[8];               // the same region of every rank (peer[rank] == local)
    bool opened[8];
};

namespace eco {
namespace {

constexpr int N_PARAMS = ECO_MPNN_N_PARAMS;
constexpr int MAX_RANKS = 8;
constexpr size_t SLOT_BYTES = ((size_t)N_PARAMS * 4 + 255) & ~size_t(255);
constexpr size_t OFF_FLAGS = 2 * SLOT_BYTES;
constexpr size_t OFF_COUNTER = OFF_FLAGS + 256;
constexpr size_t OFF_EPOCH = OFF_COUNTER + 256;        // exchanges done so far (int32): its own counter, so the caller may
constexpr size_t REGION_BYTES = OFF_EPOCH + 256;       // rewind the Adam step (e.g. after warm-up runs) without confusing peers

struct ParamTable { float* p[12]; int off[13]; };

int fill_table(const eco_mpnn_t* w, ParamTable& t) {
    const int counts[12] = {64 * 7, 63 * 8, 64 * 64, 64 * 128, 64 * 128, 64 * 128, 64 * 128, 64 * 128, 64 * 128, 64 * 64, 128, 1};
    float* ptrs[12] = {(float*)w->w_init, (float*)w->w_edge, (float*)w->w_edge_feat, (float*)w->w_msg[0], (float*)w->w_upd[0],
                       (float*)w->w_msg[1], (float*)w->w_upd[1], (float*)w->w_msg[2], (float*)w->w_upd[2], (float*)w->w_pool,
                       (float*)w->w_read, (float*)w->b_read};
    int off = 0;
    for (int k = 0; k < 12; ++k) {
        if (!ptrs[k]) return ECO_ERR_INVALID;
        t.p[k] = ptrs[k];
        t.off[k] = off;
        off += counts[k];
    }
    t.off[12] = off;
    return off == N_PARAMS ? ECO_OK : ECO_ERR_INVALID;
}

// torch.optim.Adam for element i, step t (1-based); the two bias corrections come in as 1 / (1 - beta1^t), sqrt(1 - beta2^t)
__device__ __forceinline__ void adam_element(const ParamTable& t, int i, float g, float* __restrict__ m, float* __restrict__ v,
                                             float lr_bc1, float bc2_sqrt, float beta1, float beta2, float eps, float wd) {
    int k = 0;
#pragma unroll
    for (int j = 1; j < 12; ++j) k += i >= t.off[j];
    float* p = t.p[k] + (i - t.off[k]);
    const float w = *p;
    g = fmaf(wd, w, g);
    const float mi = fmaf(beta1, m[i], (1.f - beta1) * g);
    const float vi = fmaf(beta2, v[i], (1.f - beta2) * g * g);
    m[i] = mi;
    v[i] = vi;
    *p = w - lr_bc1 * (mi / (sqrtf(vi) / bc2_sqrt + eps));
}

__device__ __forceinline__ void bias_corrections(int step, float lr, float beta1, float beta2, float& lr_bc1, float& bc2_sqrt) {
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    lr_bc1 = (float)((double)lr / bc1);
    bc2_sqrt = (float)sqrt(bc2);
}

__global__ void __launch_bounds__(256)
k_adam_dev(const ParamTable t, const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
           const int32_t* __restrict__ step_dev, const float* __restrict__ lr_dev, float beta1, float beta2, float eps, float wd,
           float grad_scale) {
    __shared__ float s_bc[2];
    if (threadIdx.x == 0) bias_corrections(*step_dev + 1, *lr_dev, beta1, beta2, s_bc[0], s_bc[1]);
    __syncthreads();
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < N_PARAMS) adam_element(t, i, grad[i] * grad_scale, m, v, s_bc[0], s_bc[1], beta1, beta2, eps, wd);
}
__global__ void k_step_inc(int32_t* step_dev, int32_t* epoch_dev) {
    *step_dev += 1;
    if (epoch_dev) *epoch_dev += 1;
}

__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct PeerTable { unsigned char* base[MAX_RANKS]; };

// One launch per update on every rank; grid <= number of SMs (all CTAs resident: they wait on one another and on the peers).
__global__ void __launch_bounds__(256)
k_dp_adam(const ParamTable t, const PeerTable peers, const int world, const int rank, const float* __restrict__ grad,
          float* __restrict__ m, float* __restrict__ v, int32_t* step_dev, const float* __restrict__ lr_dev, float beta1,
          float beta2, float eps, float wd, int* err_dev) {
    __shared__ float s_bc[2];
    unsigned char* local = peers.base[rank];
    const int epoch = *reinterpret_cast<const int*>(local + OFF_EPOCH) + 1;   // 1, 2, ...: the same on every rank
    float* slot = reinterpret_cast<float*>(local + (size_t)(epoch & 1) * SLOT_BYTES);
    int* flags = reinterpret_cast<int*>(local + OFF_FLAGS);
    unsigned long long* counter = reinterpret_cast<unsigned long long*>(local + OFF_COUNTER);
    // (a) publish this rank's gradient
    for (int i = blockIdx.x * 256 + threadIdx.x; i < N_PARAMS; i += gridDim.x * 256) slot[i] = grad[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(counter, 1ull);
        bias_corrections(*step_dev + 1, *lr_dev, beta1, beta2, s_bc[0], s_bc[1]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const unsigned long long want = (unsigned long long)gridDim.x * (unsigned long long)epoch;
        const long long t0 = clock64();
        while (atomicAdd(counter, 0ull) < want) {
            if (clock64() - t0 > 4000000000ll) { *err_dev = 1; __trap(); }
        }
        __threadfence_system();
        for (int r = 0; r < world; ++r)                    // (b) tell every rank (this one included): slot `epoch` is ready
            st_release_sys(reinterpret_cast<int*>(peers.base[r] + OFF_FLAGS) + rank, epoch);
    }
    if (threadIdx.x < world) {                             // wait for every rank's gradient of this update
        const long long t0 = clock64();
        while (ld_acquire_sys(flags + threadIdx.x) < epoch) {
            if (clock64() - t0 > 8000000000ll) { *err_dev = 2; __trap(); }
        }
    }
    __syncthreads();
    // (c) + (d): sum in rank order straight from peer memory, mean, Adam on the local parameters
    const float inv_w = 1.f / (float)world;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < N_PARAMS; i += gridDim.x * 256) {
        float g = 0.f;
        for (int r = 0; r < world; ++r)
            g += __ldcv(reinterpret_cast<const float*>(peers.base[r] + (size_t)(epoch & 1) * SLOT_BYTES) + i);
        adam_element(t, i, g * inv_w, m, v, s_bc[0], s_bc[1], beta1, beta2, eps, wd);
    }
}

}  // namespace
}  // namespace eco

using namespace eco;

extern "C" {

int eco_mpnn_adam_dev(const eco_mpnn_t* w, const float* grad, float* exp_avg, float* exp_avg_sq, int32_t* step_dev,
                      const float* lr_dev, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                      void* stream) {
    ECO_CHECK_ARG(w && grad && exp_avg && exp_avg_sq && step_dev && lr_dev, ECO_ERR_INVALID, "eco_mpnn_adam_dev: null argument");
    ParamTable t;
    ECO_CHECK_ARG(fill_table(w, t) == ECO_OK, ECO_ERR_INVALID, "eco_mpnn_adam_dev: weight pointers missing");
    cudaStream_t st = (cudaStream_t)stream;
    k_adam_dev<<<(N_PARAMS + 255) / 256, 256, 0, st>>>(t, grad, exp_avg, exp_avg_sq, step_dev, lr_dev, beta1, beta2, eps,
                                                       weight_decay, grad_scale);
    ECO_LAUNCH_CHECK();
    k_step_inc<<<1, 1, 0, st>>>(step_dev, nullptr);
    ECO_LAUNCH_CHECK();
    return ECO_OK;
}

int eco_dp_create(eco_dp_t** out, int32_t world, int32_t rank) {
    ECO_CHECK_ARG(out && world >= 1 && world <= MAX_RANKS && rank >= 0 && rank < world, ECO_ERR_INVALID,
                  "eco_dp_create: world must be 1..%d and 0 <= rank < world", MAX_RANKS);
    eco_dp* dp = new (std::nothrow) eco_dp();
    ECO_CHECK_ARG(dp, ECO_ERR_INVALID, "eco_dp_create: out of memory");
    dp->world = world;
    dp->rank = rank;
    dp->n_sm = device_sm_count();
    for (int r = 0; r < 8; ++r) { dp->peer[r] = nullptr; dp->opened[r] = false; }
    if (cudaMalloc(&dp->local, REGION_BYTES) != cudaSuccess || cudaMemset(dp->local, 0, REGION_BYTES) != cudaSuccess) {
        set_error("eco_dp_create: cudaMalloc of the exchange region failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete dp;
        return ECO_ERR_CUDA;
    }
    dp->peer[rank] = dp->local;
    *out = dp;
    return ECO_OK;
}

int eco_dp_handle(const eco_dp_t* dp, void* handle64) {
    ECO_CHECK_ARG(dp && handle64, ECO_ERR_INVALID, "eco_dp_handle: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == ECO_DP_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    ECO_CUDA(cudaIpcGetMemHandle(&h, dp->local));
    memcpy(handle64, &h, sizeof(h));
    return ECO_OK;
}

int eco_dp_open(eco_dp_t* dp, const void* handles) {
    ECO_CHECK_ARG(dp && handles, ECO_ERR_INVALID, "eco_dp_open: null argument");
    for (int r = 0; r < dp->world; ++r) {
        if (r == dp->rank || dp->peer[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const unsigned char*)handles + (size_t)r * ECO_DP_HANDLE_BYTES, sizeof(h));
        void* p = nullptr;
        ECO_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        dp->peer[r] = (unsigned char*)p;
        dp->opened[r] = true;
    }
    return ECO_OK;
}

int eco_dp_adam(eco_dp_t* dp, const eco_mpnn_t* w, const float* grad, float* exp_avg, float* exp_avg_sq, int32_t* step_dev,
                const float* lr_dev, float beta1, float beta2, float eps, float weight_decay, int32_t* err_dev, void* stream) {
    ECO_CHECK_ARG(dp && w && grad && exp_avg && exp_avg_sq && step_dev && lr_dev && err_dev, ECO_ERR_INVALID,
                  "eco_dp_adam: null argument");
    ParamTable t;
    ECO_CHECK_ARG(fill_table(w, t) == ECO_OK, ECO_ERR_INVALID, "eco_dp_adam: weight pointers missing");
    PeerTable pt;
    for (int r = 0; r < MAX_RANKS; ++r) pt.base[r] = r < dp->world ? dp->peer[r] : nullptr;
    for (int r = 0; r < dp->world; ++r)
        ECO_CHECK_ARG(pt.base[r], ECO_ERR_INVALID, "eco_dp_adam: peer %d not opened (eco_dp_open)", r);
    cudaStream_t st = (cudaStream_t)stream;
    int grid = (N_PARAMS + 255) / 256;
    if (grid > dp->n_sm) grid = dp->n_sm;             // every CTA must be resident: they wait on one another
    k_dp_adam<<<grid, 256, 0, st>>>(t, pt, dp->world, dp->rank, grad, exp_avg, exp_avg_sq, step_dev, lr_dev, beta1, beta2, eps,
                                    weight_decay, err_dev);
    ECO_LAUNCH_CHECK();
    k_step_inc<<<1, 1, 0, st>>>(step_dev, reinterpret_cast<int32_t*>(dp->local + OFF_EPOCH));
    ECO_LAUNCH_CHECK();
    return ECO_OK;
}

void eco_dp_destroy(eco_dp_t* dp) {
    if (!dp) return;
    for (int r = 0; r < 8; ++r)
        if (dp->opened[r] && dp->peer[r]) cudaIpcCloseMemHandle(dp->peer[r]);
    if (dp->local) cudaFree(dp->local);
    delete dp;
}

}  // extern "C"
