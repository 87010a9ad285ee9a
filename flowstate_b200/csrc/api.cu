// Error plumbing of the C ABI.
#include <stdarg.h>

#include "common.cuh"

namespace fs {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_check(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return FS_OK;
    set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
    return FS_ERR_CUDA;
}
}  // namespace fs

namespace fs {
static unsigned long long g_launches = 0;   // single host thread per process
void count_launch(int n) { g_launches += (unsigned long long)n; }
}  // namespace fs

extern "C" unsigned long long fs_launch_count(void) { return fs::g_launches; }
extern "C" const char* fs_last_error(void) { return fs::g_err; }
extern "C" int fs_version(void) { return 100; }

// The library links its own (static) CUDA runtime, whose "current device" is separate from the
// caller's runtime: callers on a multi-GPU box bind it to the device that owns their buffers.
extern "C" int fs_set_device(int device) {
    return fs::cuda_check(cudaSetDevice(device), "cudaSetDevice");
}
