// abi.cu -- error plumbing and the view-table upload of the C ABI (include/tomo_b200.h).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>
#include "tomo_common.h"

static thread_local char g_err[512] = "";

extern "C" void tomo_set_error(const char* msg)
{
    std::snprintf(g_err, sizeof(g_err), "%s", msg ? msg : "");
}

int tomo_check_cuda(cudaError_t e, const char* what)
{
    if (e == cudaSuccess) return 0;
    std::snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return (int)e;
}

extern "C" int tomo_version(void) { return TOMO_B200_VERSION; }

extern "C" const char* tomo_last_error(void) { return g_err; }

extern "C" size_t tomo_views_bytes(int n_proj)
{
    return n_proj > 0 ? sizeof(double) * TOMO_VIEW_STRIDE * (size_t)n_proj : 0;
}

extern "C" int tomo_views_kinds(const double* views_host, int n_proj)
{
    if (!views_host || n_proj <= 0) return 0;
    int kinds = TOMO_KINDS_KNOWN;
    for (int v = 0; v < n_proj; ++v) {
        const double* V = views_host + (size_t)v * TOMO_VIEW_STRIDE;
        const bool sep = V[V_SEP] != 0.0, col = V[V_NCOL] != 0.0;
        kinds |= sep ? TOMO_KINDS_SEPARABLE : (V[V_ZQ] != 0.0 ? TOMO_KINDS_ZQUAD : TOMO_KINDS_GENERIC);
        if (!col) kinds |= TOMO_KINDS_UNCOLOURED;
        else if (!sep) kinds |= TOMO_KINDS_TILE;
    }
    return kinds;
}

extern "C" int tomo_views_upload(const TomoGeom* g, const double* poses, int n_proj, void* views_dev, void* stream)
{
    if (!views_dev) { tomo_set_error("tomo_views_upload: views_dev is NULL"); return TOMO_E_ARG; }
    if (n_proj <= 0) { tomo_set_error("tomo_views_upload: n_proj <= 0"); return TOMO_E_ARG; }
    std::vector<double> host((size_t)n_proj * TOMO_VIEW_STRIDE);
    if (int e = tomo_views_compute_host(g, poses, n_proj, host.data())) return e;
    // pageable source: the runtime stages the buffer before returning, so `host` may die here
    cudaError_t ce = cudaMemcpyAsync(views_dev, host.data(), tomo_views_bytes(n_proj), cudaMemcpyHostToDevice,
                                     (cudaStream_t)stream);
    return tomo_check_cuda(ce, "tomo_views_upload: cudaMemcpyAsync");
}
