// Kernel family 2b placeholder: tcgen05 MPNN forward (filled in by the next milestone).
#include "eco_common.cuh"

namespace eco {
bool mpnn_tc_supported(const eco_graphs_t*) { return false; }
size_t mpnn_tc_scratch_bytes(int, int) { return 0; }
size_t mpnn_tc_packed_bytes() { return 0; }
int launch_mpnn_pack(const eco_mpnn_t*, void*, cudaStream_t) {
    set_error("tcgen05 MPNN path not built");
    return ECO_ERR_UNSUPPORTED;
}
int launch_mpnn_tc(const eco_graphs_t*, const eco_mpnn_t*, int, const int32_t*, const float*, const float*, float,
                   float*, int32_t*, void*, cudaStream_t) {
    set_error("tcgen05 MPNN path not built");
    return ECO_ERR_UNSUPPORTED;
}
}  // namespace eco
