// Internal (C++) entry points shared between translation units of the library.
#pragma once
#include "common.cuh"

namespace ds {

// spmm.cu
int spmm_km(const int32_t* brow, const int32_t* bcol, int64_t n_nodes, const double* Kval, const double* Mblk,
            double shift, const double* X, int64_t ldx, int ncols, double alpha, double beta, const double* Y0,
            int64_t ldy0, double* Y, int64_t ldy, cudaStream_t stream);
int spmm_dual(const int32_t* brow, const int32_t* bcol, int64_t n_nodes, const double* Kval, const double* Mblk,
              const double* X, int64_t ldx, int ncols, double* YK, int64_t ldyk, double* YM, int64_t ldym,
              cudaStream_t stream);
int block_jacobi(const int32_t* brow, const int32_t* bcol, int64_t n_nodes, const double* Kval, const double* Mblk,
                 double shift, double* invD, cudaStream_t stream);
// `degree` block-Jacobi Chebyshev steps on K + shift*M applied to R; result points at Z0 or Z1.
int cheb_precond(const int32_t* brow, const int32_t* bcol, int64_t n_nodes, const double* Kval, const double* Mblk,
                 double shift, const double* invD, double lmin, double lmax, int degree, const double* R, int64_t ldr,
                 int ncols, double* Z0, double* Z1, int64_t ldz, double** result, cudaStream_t stream);

// dense.cu
int64_t gram_scratch_elems(int p, int q);
int gram_f64(const double* A, int64_t lda, int p, const double* B, int64_t ldb, int q, int64_t n, double* G,
             int64_t ldg, double* partial, cudaStream_t stream);
// Y = beta Y + alpha A C
int block_gemm_f64(const double* A, int64_t lda, int p, const double* C, int64_t ldc, int q, int64_t n, double alpha,
                   double beta, double* Y, int64_t ldy, cudaStream_t stream);
// idx_host (may be NULL = identity): slot of compact index i inside the ldg x ldg Gram storage;
// only the upper triangle of the storage is read; rows of C are written at the mapped slots.
// sigma < 0 selects an automatic shift |sigma| * mean(diag(scaled GK)).
int eigh_generalized_f64(const double* GK, const double* GM, int N, int64_t ldg, const int* idx_host, double sigma,
                         double* theta, double* C, int64_t ldc, double* scratch, int* info, cudaStream_t stream);

}  // namespace ds
