// C-ABI entry points of the pieces a row-partitioned (multi-GPU) LOBPCG step is assembled from.
//
// One large mesh on several GPUs (BASELINE.json configs[2]): rank r owns a contiguous slab of
// node rows of K, M and of the iterate blocks.  The host driver is diffsound_b200/parallel/rowpart_lobpcg.py (Python, like
// the reference's own LOBPCG, /root/reference/src/lobpcg/_lobpcg.py:344-477); it strings these kernels together with the
// exchanges of the path: NCCL all-reduce of the Gram strips / residual norms / partial coarse residuals and an all-gather of
// the new fp32 search block.  Everything here is a thin wrapper over kernels the single-GPU driver (csrc/lobpcg.cu) uses.
#include "common.cuh"
#include "../../include/diffsound_sm100.h"
#include "kernels.cuh"

namespace ds {

// rc[I] += 0.5 sum over the fine nodes of I's gather list that lie in [fine_lo, fine_hi); res holds the rows of that range
__global__ void k_restrict32_range(const int32_t* __restrict__ rptr, const int32_t* __restrict__ rlist, int64_t n_coarse,
                                   const float* __restrict__ res, int c4x3, float* __restrict__ rc, int fine_lo, int fine_hi) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_coarse * c4x3) return;
    const int64_t I = t / c4x3;
    const int q = (int)(t - I * c4x3);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    const int e = rptr[I + 1];
    for (int u = rptr[I]; u < e; ++u) {
        const int f = rlist[u];
        if (f < fine_lo || f >= fine_hi) continue;
        const float4 v = __ldg(reinterpret_cast<const float4*>(res) + (int64_t)(f - fine_lo) * c4x3 + q);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    s.x *= 0.5f; s.y *= 0.5f; s.z *= 0.5f; s.w *= 0.5f;
    reinterpret_cast<float4*>(rc)[t] = s;
}

}  // namespace ds

using namespace ds;

extern "C" int ds_gather_cols_f32(const double* src, int64_t lds, const int* cols_host, int count, int width, int64_t n,
                                  float* dst, void* stream) {
    DS_REQUIRE(src && dst && cols_host && count >= 0 && count <= width && width <= 128 && width % 4 == 0, "ds_gather_cols_f32: bad argument");
    ColIdx ci;
    for (int s = 0; s < 128; ++s) ci.v[s] = (short)(s < count ? cols_host[s] : 0);
    return gather_cols_f32(src, lds, ci, count, width, n, dst, (cudaStream_t)stream, nullptr);
}

extern "C" int ds_widen_f32(const float* src, int width, int64_t n, double* dst, int64_t ldd, void* stream) {
    DS_REQUIRE(src && dst, "ds_widen_f32: null argument");
    return widen_f32(src, width, n, dst, ldd, (cudaStream_t)stream, nullptr);
}

extern "C" int ds_jacobi32(const float* invD, const float* R, int64_t n_nodes, int ncols, double cc, float* Out, void* stream) {
    DS_REQUIRE(invD && R && Out && ncols % 4 == 0, "ds_jacobi32: bad argument");
    return jacobi32(invD, R, n_nodes, ncols, (float)cc, Out, (cudaStream_t)stream);
}

extern "C" int ds_spmm_dual_z32(const int32_t* brow, const int32_t* bcolP, int64_t n_rows, int64_t nnzb, const int32_t* chunk_row,
                                const double* Kval, const double* Mblk, const float* Z, int ncols, double* YK, int64_t ldyk,
                                double* YM, int64_t ldym, void* stream) {
    DS_REQUIRE(brow && bcolP && chunk_row && Kval && Mblk && Z && YK && YM, "ds_spmm_dual_z32: null argument");
    return spmm_dual_z32(brow, brow, bcolP, nullptr, chunk_row, spmm32_chunk_count(n_rows), n_rows, Kval, Mblk, Z, ncols, YK, ldyk,
                         YM, ldym, (cudaStream_t)stream, nnzb);
}

extern "C" int ds_pmg_restrict32_range(const int32_t* rptr, const int32_t* rlist, int64_t n_coarse, const float* res_local,
                                       int ncols, int64_t fine_lo, int64_t fine_hi, float* rc, void* stream) {
    DS_REQUIRE(rptr && rlist && res_local && rc && ncols % 4 == 0 && fine_lo >= 0 && fine_hi >= fine_lo, "ds_pmg_restrict32_range: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(PROF_TRANSFER, st);
    const int q = 3 * ncols / 4;
    k_restrict32_range<<<(unsigned)ceil_div(n_coarse * q, 256), 256, 0, st>>>(rptr, rlist, n_coarse, res_local, q, rc, (int)fine_lo,
                                                                             (int)fine_hi);
    DS_LAUNCH_CHECK();
    return DS_OK;
}

extern "C" int ds_pmg_prolong64(const int32_t* parents, int64_t n_fine, const double* xc, int64_t ldc, int w, double* x, int64_t ldx,
                                void* stream) {
    DS_REQUIRE(parents && xc && x && ((uintptr_t)parents % 8) == 0, "ds_pmg_prolong64: bad argument (parents must be 8-byte aligned)");
    return prolong64(parents, n_fine, xc, ldc, w, x, ldx, (cudaStream_t)stream);
}

extern "C" int ds_gram_insert_f64(double* GK, double* GM, int64_t ldg, const double* GsK, const double* GsM, int64_t lds, int m, int wa,
                                  void* stream) {
    return gram_insert(GK, GM, ldg, GsK, GsM, lds, m, wa, (cudaStream_t)stream);
}

extern "C" int ds_sym_upper_f64(double* GK, double* GM, int64_t ldg, int N, void* stream) {
    DS_REQUIRE(GK && GM && N > 0 && ldg >= N, "ds_sym_upper_f64: bad argument");
    return sym_upper(GK, GM, ldg, N, (cudaStream_t)stream);
}

extern "C" int ds_eigh_generalized_idx_f64(const double* GK, const double* GM, int N, int64_t ldg, const int* idx_host, double sigma,
                                           double* theta, double* C, int64_t ldc, double* scratch, int* info, void* stream) {
    DS_REQUIRE(idx_host, "ds_eigh_generalized_idx_f64: idx_host is NULL");
    return eigh_generalized_f64(GK, GM, N, ldg, idx_host, sigma, theta, C, ldc, scratch, info, (cudaStream_t)stream);
}
