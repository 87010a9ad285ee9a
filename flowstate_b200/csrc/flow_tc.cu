// Tensor-core (tcgen05 / TMEM / TMA) conditioner - placeholder until the fused kernel lands.
#include "flow.cuh"

namespace fs {
int tc_pack(fs_flow* f, const fs_flow_desc*) {
    f->tc = nullptr;
    return FS_OK;
}
void tc_free(fs_flow*) {}
size_t tc_workspace_bytes(const fs_flow*, int) { return 0; }
int tc_conditioner(fs_flow*, int, const float*, int, float*, void*, size_t, cudaStream_t) {
    set_error("tensor-core conditioner not available");
    return FS_ERR_UNSUPPORTED;
}
}  // namespace fs
