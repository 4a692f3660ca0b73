// common.cuh -- shared helpers of libdpiso (error plumbing, index helpers, reductions).
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <math.h>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define DPISO_HD __host__ __device__ __forceinline__
#else
#define DPISO_HD inline
#endif

#include "../../include/dpiso.h"

namespace dpiso {

void set_error(const char *fmt, ...);

#ifdef __CUDACC__
#define DPISO_CUDA_TRY(expr)                                                                           \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            dpiso::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return DPISO_ECUDA;                                                                        \
        }                                                                                              \
    } while (0)

#define DPISO_CHECK_LAUNCH()                                                                           \
    do {                                                                                               \
        cudaError_t _e = cudaGetLastError();                                                           \
        if (_e != cudaSuccess) {                                                                       \
            dpiso::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return DPISO_ECUDA;                                                                        \
        }                                                                                              \
    } while (0)
#endif

#define DPISO_REQUIRE(cond, ...)           \
    do {                                   \
        if (!(cond)) {                     \
            dpiso::set_error(__VA_ARGS__); \
            return DPISO_EINVAL;           \
        }                                  \
    } while (0)

DPISO_HD int wrapi(int i, int n) {
    i %= n;
    return i < 0 ? i + n : i;
}
DPISO_HD int clampi(int i, int lo, int hi) { return i < lo ? lo : (i > hi ? hi : i); }

// geometry of one sample
struct Grid {
    int ny, nx;        // centred resolution
    int per_x, per_y;  // periodic flags
    int n_u, n_v;      // face counts: ny*(nx+1), (ny+1)*nx
    int nnz_u, nnz_v;  // CSR entries per component
    DPISO_HD int nf() const { return n_u + n_v; }
    DPISO_HD int nc() const { return ny * nx; }
    DPISO_HD int nnz() const { return nnz_u + nnz_v; }
};

inline Grid make_grid(int ny, int nx, int per_x, int per_y) {
    Grid g;
    g.ny = ny; g.nx = nx; g.per_x = per_x ? 1 : 0; g.per_y = per_y ? 1 : 0;
    g.n_u = ny * (nx + 1);
    g.n_v = (ny + 1) * nx;
    // diffpiso/piso_tf.py:102-106
    g.nnz_u = 5 * g.n_u - 2 * ny * (1 - g.per_x) - 2 * (nx + 1) * (1 - g.per_y);
    g.nnz_v = 5 * g.n_v - 2 * (ny + 1) * (1 - g.per_x) - 2 * nx * (1 - g.per_y);
    return g;
}

}  // namespace dpiso
