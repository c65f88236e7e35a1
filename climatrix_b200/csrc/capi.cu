// Host-side bookkeeping and the small exported entry points of libclimatrix_b200.
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace cmx {
namespace {
thread_local long long g_launches = 0;
thread_local char g_cuda_err[256] = "";
}  // namespace

void note_launch() { ++g_launches; }

int cuda_fail(cudaError_t e) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
    return CMX_ERR_CUDA;
}

int check_launch() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? CMX_OK : cuda_fail(e);
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        return 148;  // no device (build container): the B200 figure, only used for size queries
    }
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
            cudaGetLastError();
            n = 148;
        }
        cached = n;
        cached_dev = dev;
    }
    return cached;
}
}  // namespace cmx

extern "C" {

int cmx_version(void) { return 100; }  // 0.1.0

const char* cmx_strerror(int code) {
    switch (code) {
        case CMX_OK: return "ok";
        case CMX_ERR_BAD_ARG: return "bad argument";
        case CMX_ERR_K_GT_N: return "k exceeds the number of samples";
        case CMX_ERR_K_TOO_LARGE: return "k exceeds CMX_MAX_K";
        case CMX_ERR_WORKSPACE: return "workspace missing or too small";
        case CMX_ERR_CUDA: return "CUDA error";
        case CMX_ERR_UNSUPPORTED: return "unsupported option";
        case CMX_ERR_TOO_MANY: return "too many samples or targets";
        default: return "unknown error";
    }
}

const char* cmx_last_cuda_error(void) { return cmx::g_cuda_err; }

int64_t cmx_launch_count(void) { return cmx::g_launches; }
void cmx_reset_launch_count(void) { cmx::g_launches = 0; }

size_t cmx_knn_workspace_bytes(int k) { return cmx::knn_workspace_bytes(k); }

int cmx_knn(const double* s_a, const double* s_b, int64_t n, const double* t_a, int64_t n_a, const double* t_b,
            int64_t n_b, int grid, int k, int32_t* idx_out, double* dist_out, void* ws, size_t ws_bytes,
            void* stream) {
    if (!idx_out) return CMX_ERR_BAD_ARG;
    cmx::KnnJob job{};
    job.s_a = s_a;
    job.s_b = s_b;
    job.n = n;
    job.t_a = t_a;
    job.t_b = t_b;
    job.n_a = n_a;
    job.n_b = n_b;
    job.grid = grid;
    job.k = k;
    job.idx_out = idx_out;
    job.dist_out = dist_out;
    job.workspace = ws;
    job.workspace_bytes = ws_bytes;
    return cmx::run_knn(job, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
