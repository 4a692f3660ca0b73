// oracle/ref_shim.cu -- TEST INFRASTRUCTURE.  extern "C" entry points around the REFERENCE's own CUDA launchers
// (compiled in place from /root/reference/CUDAsrc by oracle/build.py --ref into oracle/_ref/libdiffpiso_ref.so).
// Used only by tests/test_gpu_reference_pin.py on the GPU box to pin the CPU oracle against the reference kernels:
//   CentralDifferenceMatrixCsrKernelLauncher   central_difference_csr_op.cu.cc:543-664
//   LaplaceMatrixKernelLauncher                 laplace_op.cu.cc:191-239
//   LaunchPressureKernel (float / double)       pressure_solve_op.cu.cc:140-696
// All arguments are HOST arrays; the shim stages them on the device exactly like the TF op shims hand them over
// (central_difference_csr_op.cc:50-104, pressure_solve_op.cc:108-234).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

void CentralDifferenceMatrixCsrKernelLauncher(const float *velocity, float *const csrMatVal, int *const csrColInd,
                                              int *const csrRowPtr, float *const diagonalArray,
                                              const bool *boolDirichletMask, const float *active_mask,
                                              const float *accessible_mask, const float *viscosity,
                                              const int *dimensions, const int *padDepth, const float *cellArea,
                                              const float *gridSpacing, const bool boolViscosityField, const int dimSize,
                                              const bool *noSlipWall, const bool *boolPeriodic, const float *beta,
                                              const int unrolling_step);
void LaplaceMatrixKernelLauncher(const int *dimensions, const int dim_size, const int dim_product,
                                 const float *active_mask, const float *fluid_mask, const int *mask_dimensions,
                                 float *laplace_matrix, int *cords, const float *advection_influence,
                                 const int *staggered_dimensions);
void LaplaceMatrixKernelLauncher(const int *dimensions, const int dim_size, const int dim_product,
                                 const float *active_mask, const float *fluid_mask, const int *mask_dimensions,
                                 double *laplace_matrix, int *cords, const float *advection_influence,
                                 const int *staggered_dimensions);
void LaunchPressureKernel(const int *dimensions, const int dim_product, const int dim_size, const float *laplace_matrix,
                          float *p, float *z, float *r, float *divergence, float *x, bool *threshold_reached,
                          const float *accuracy_ptr, const int *max_iterations_ptr, const int batch_size,
                          int *iterations_gpu, const bool *boolPeriodic, const bool *laplace_rank_deficient,
                          const bool init_with_zeros, const int residual_reset_steps, const int randomized_restarts,
                          const int unrolling_step);
void LaunchPressureKernel(const int *dimensions, const int dim_product, const int dim_size, const double *laplace_matrix,
                          double *p, double *z, double *r, double *divergence, double *x, bool *threshold_reached,
                          const float *accuracy_ptr, const int *max_iterations_ptr, const int batch_size,
                          int *iterations_gpu, const bool *boolPeriodic, const bool *laplace_rank_deficient,
                          const bool init_with_zeros, const int residual_reset_steps, const int randomized_restarts,
                          const int unrolling_step);

namespace {
struct Arena {
    std::vector<void *> ptrs;
    template <typename T> T *up(const T *host, size_t n) {
        T *d = nullptr;
        cudaMalloc((void **)&d, (n ? n : 1) * sizeof(T));
        if (host) cudaMemcpy(d, host, n * sizeof(T), cudaMemcpyHostToDevice);
        else cudaMemset(d, 0, (n ? n : 1) * sizeof(T));
        ptrs.push_back(d);
        return d;
    }
    ~Arena() { for (void *p : ptrs) cudaFree(p); }
};
template <typename T> void down(T *host, const T *dev, size_t n) { cudaMemcpy(host, dev, n * sizeof(T), cudaMemcpyDeviceToHost); }
}  // namespace

extern "C" {

// dims4 = [Nx+1, Ny, Nx, Ny+1]; periodic_xy = (x, y); cell_area / grid_spacing as built by piso_tf.py:96-97
int ref_assemble(int n_pad_total, const float *velocity_padded, int n_u, int n_v, int nnz, const unsigned char *dirichlet,
                 int n_mask, const float *active, const float *access, const float *visc, int n_visc, const int *dims4,
                 const float *cell_area, const float *grid_spacing, const unsigned char *noslip,
                 const unsigned char *periodic_xy, float beta, float *values, int *col_ind, int *row_ptr, float *a_diag) {
    Arena A;
    const int pad4[4] = {1, 1, 1, 1};
    float *d_vel = A.up(velocity_padded, n_pad_total);
    float *d_val = A.up<float>(nullptr, nnz);
    int *d_ci = A.up<int>(nullptr, nnz);
    int *d_rp = A.up<int>(nullptr, n_u + n_v + 2);
    float *d_diag = A.up<float>(nullptr, n_u + n_v);
    bool *d_dir = (bool *)A.up(dirichlet, n_u + n_v);
    float *d_act = A.up(active, n_mask), *d_acc = A.up(access, n_mask);
    float *d_visc = A.up(visc, n_visc);
    int *d_dims = A.up(dims4, 4), *d_pad = A.up(pad4, 4);
    float *d_ca = A.up(cell_area, 2), *d_gs = A.up(grid_spacing, 2);
    bool *d_ns = (bool *)A.up(noslip, n_mask);
    bool *d_per = (bool *)A.up(periodic_xy, 2);
    float *d_beta = A.up(&beta, 1);
    CentralDifferenceMatrixCsrKernelLauncher(d_vel, d_val, d_ci, d_rp, d_diag, d_dir, d_act, d_acc, d_visc, d_dims, d_pad,
                                             d_ca, d_gs, n_visc > 1, 2, d_ns, d_per, d_beta, 0);
    cudaDeviceSynchronize();
    down(values, d_val, nnz); down(col_ind, d_ci, nnz); down(row_ptr, d_rp, n_u + n_v + 2); down(a_diag, d_diag, n_u + n_v);
    return (int)cudaGetLastError();
}

// dims_xy = (Nx, Ny); k_faces flattened [v, u]; fp64 != 0 selects the double path.  lap_out / x_out are T arrays.
int ref_pressure_solve(int fp64, int nx, int ny, int batch, const float *active, const float *fluid, const float *k_faces,
                       const void *divergence, float accuracy, int max_it, const unsigned char *periodic_xy,
                       int rank_deficient, int residual_reset, void *lap_out, void *x_out, int *iterations) {
    Arena A;
    const int dims[2] = {nx, ny}, mdims[2] = {nx + 2, ny + 2}, sdims[4] = {nx + 1, ny, nx, ny + 1};
    const int nc = nx * ny, n_mask = (nx + 2) * (ny + 2), n_faces = (nx + 1) * ny + nx * (ny + 1);
    int *d_dims = A.up(dims, 2), *d_mdims = A.up(mdims, 2), *d_sdims = A.up(sdims, 4);
    float *d_act = A.up(active, n_mask), *d_fl = A.up(fluid, n_mask), *d_k = A.up(k_faces, n_faces);
    int *d_cords = A.up<int>(nullptr, (size_t)nc * 2);
    bool *d_thr = A.up<bool>(nullptr, batch);
    float *d_acc = A.up(&accuracy, 1);
    int *d_maxit = A.up(&max_it, 1);
    int *d_it = A.up<int>(nullptr, 1);
    bool *d_per = (bool *)A.up(periodic_xy, 2);
    const bool rd = rank_deficient != 0;
    bool *d_rd = A.up(&rd, 1);
    if (fp64) {
        double *d_lap = A.up<double>(nullptr, (size_t)nc * 5);
        double *d_div = A.up((const double *)divergence, (size_t)nc * batch);
        double *d_p = A.up<double>(nullptr, (size_t)nc * batch), *d_r = A.up<double>(nullptr, (size_t)nc * batch);
        double *d_z = A.up<double>(nullptr, (size_t)nc * batch), *d_x = A.up<double>(nullptr, (size_t)nc * batch);
        LaplaceMatrixKernelLauncher(d_dims, 2, nc, d_act, d_fl, d_mdims, d_lap, d_cords, d_k, d_sdims);
        LaunchPressureKernel(d_dims, nc, 2, d_lap, d_p, d_z, d_r, d_div, d_x, d_thr, d_acc, d_maxit, batch, d_it, d_per,
                             d_rd, true, residual_reset, 0, 0);
        cudaDeviceSynchronize();
        down((double *)lap_out, d_lap, (size_t)nc * 5); down((double *)x_out, d_x, (size_t)nc * batch);
    } else {
        float *d_lap = A.up<float>(nullptr, (size_t)nc * 5);
        float *d_div = A.up((const float *)divergence, (size_t)nc * batch);
        float *d_p = A.up<float>(nullptr, (size_t)nc * batch), *d_r = A.up<float>(nullptr, (size_t)nc * batch);
        float *d_z = A.up<float>(nullptr, (size_t)nc * batch), *d_x = A.up<float>(nullptr, (size_t)nc * batch);
        LaplaceMatrixKernelLauncher(d_dims, 2, nc, d_act, d_fl, d_mdims, d_lap, d_cords, d_k, d_sdims);
        LaunchPressureKernel(d_dims, nc, 2, d_lap, d_p, d_z, d_r, d_div, d_x, d_thr, d_acc, d_maxit, batch, d_it, d_per,
                             d_rd, true, residual_reset, 0, 0);
        cudaDeviceSynchronize();
        down((float *)lap_out, d_lap, (size_t)nc * 5); down((float *)x_out, d_x, (size_t)nc * batch);
    }
    down(iterations, d_it, 1);
    return (int)cudaGetLastError();
}

}  // extern "C"
