// Block-CSR (3x3 node blocks) sparse x dense-block products, FP64.
//
// Reference behaviour replaced (under /root/reference/src): torch sparse COO
// `stiff_matrix @ U`, `mass_matrix @ U` (diffelastic/diff_model.py:385,395-397)
// and torch.sparse.mm inside lobpcg/_linalg_utils.py:27-39.
//
// Layout: K values in the reference's scalar-CSR (row, col) order -- row 3i+c of
// node i is the contiguous run Kval[9*brow[i] + c*3*deg .. +3*deg) -- but indexed
// through the node-level block pattern (brow, bcol), so index traffic is 4 B per
// 3x3 block instead of 4 B per scalar.  M is stored as one scalar per block
// (M = Mblk (x) I3).  Dense blocks are row-major (n x ncols), ncols = 16*CPL.
//
// Mapping: one warp per node row.  The two half-warps walk alternate neighbour
// blocks; lane cl of a half-warp owns columns cl + 16 t (t < CPL) and all three
// components of the node, so a block costs 9 broadcast K loads, 3*CPL coalesced X
// loads and 9*CPL DFMA per lane; halves are combined with one shuffle at the end.
#include "common.cuh"
#include "../../include/diffsound_sm100.h"
#include "kernels.cuh"

namespace ds {

template <int CPL, bool HAS_K, bool HAS_M, bool DUAL>
__device__ __forceinline__ void row_product(const int32_t* __restrict__ bcol, const double* __restrict__ Kval,
                                            const double* __restrict__ Mblk, double shift,
                                            const double* __restrict__ X, int64_t ldx, int64_t b0, int deg,
                                            int half, int cl, double (&acc)[3][CPL], double (&accm)[3][CPL]) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int t = 0; t < CPL; ++t) {
            acc[c][t] = 0.0;
            accm[c][t] = 0.0;
        }
    const double* kbase = HAS_K ? Kval + 9 * b0 : nullptr;
    const int64_t rs = 3 * (int64_t)deg;
    for (int p = half; p < deg; p += 2) {
        int64_t j = bcol[b0 + p];
        const double* xr = X + 3 * j * ldx + cl;
        double x[3][CPL];
#pragma unroll
        for (int d = 0; d < 3; ++d)
#pragma unroll
            for (int t = 0; t < CPL; ++t) x[d][t] = __ldg(xr + d * ldx + 16 * t);
        double k[3][3];
        if (HAS_K) {
            const double* kp = kbase + 3 * p;
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int d = 0; d < 3; ++d) k[c][d] = __ldg(kp + c * rs + d);
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int d = 0; d < 3; ++d) k[c][d] = 0.0;
        }
        double m = 0.0;
        if (HAS_M) {
            m = __ldg(Mblk + b0 + p);
            if (!DUAL) {
                k[0][0] += shift * m;
                k[1][1] += shift * m;
                k[2][2] += shift * m;
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int t = 0; t < CPL; ++t) {
                double a = acc[c][t];
                a = fma(k[c][0], x[0][t], a);
                a = fma(k[c][1], x[1][t], a);
                a = fma(k[c][2], x[2][t], a);
                acc[c][t] = a;
                if (DUAL) accm[c][t] = fma(m, x[c][t], accm[c][t]);
            }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int t = 0; t < CPL; ++t) {
            acc[c][t] += __shfl_xor_sync(0xffffffffu, acc[c][t], 16);
            if (DUAL) accm[c][t] += __shfl_xor_sync(0xffffffffu, accm[c][t], 16);
        }
}

template <int CPL, bool HAS_K, bool HAS_M>
__global__ void __launch_bounds__(256)
k_spmm(const int32_t* __restrict__ brow, const int32_t* __restrict__ bcol, int64_t n_nodes,
       const double* __restrict__ Kval, const double* __restrict__ Mblk, double shift,
       const double* __restrict__ X, int64_t ldx, double alpha, double beta, const double* __restrict__ Y0,
       int64_t ldy0, double* __restrict__ Y, int64_t ldy) {
    int lane = threadIdx.x & 31, half = lane >> 4, cl = lane & 15;
    int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= n_nodes) return;
    int64_t b0 = brow[row];
    int deg = (int)(brow[row + 1] - b0);
    double acc[3][CPL], accm[3][CPL];
    row_product<CPL, HAS_K, HAS_M, false>(bcol, Kval, Mblk, shift, X, ldx, b0, deg, half, cl, acc, accm);
    // half 0 stores component rows {0, 2(lower cols)}, half 1 stores {1, 2(upper)}: keep it simple --
    // half 0 writes c = 0,1 ; half 1 writes c = 2 plus nothing else would unbalance, so split by t parity.
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int t = 0; t < CPL; ++t) {
            if (((c * CPL + t) & 1) != half) continue;
            int64_t col = cl + 16 * t;
            double v = alpha * acc[c][t];
            if (beta != 0.0) v = fma(beta, Y0[(3 * row + c) * ldy0 + col], v);
            Y[(3 * row + c) * ldy + col] = v;
        }
}

template <int CPL>
__global__ void __launch_bounds__(256)
k_spmm_dual(const int32_t* __restrict__ brow, const int32_t* __restrict__ bcol, int64_t n_nodes,
            const double* __restrict__ Kval, const double* __restrict__ Mblk, const double* __restrict__ X,
            int64_t ldx, double* __restrict__ YK, int64_t ldyk, double* __restrict__ YM, int64_t ldym) {
    int lane = threadIdx.x & 31, half = lane >> 4, cl = lane & 15;
    int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= n_nodes) return;
    int64_t b0 = brow[row];
    int deg = (int)(brow[row + 1] - b0);
    double acc[3][CPL], accm[3][CPL];
    row_product<CPL, true, true, true>(bcol, Kval, Mblk, 0.0, X, ldx, b0, deg, half, cl, acc, accm);
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int t = 0; t < CPL; ++t) {
            int64_t col = cl + 16 * t;
            if (half == 0) YK[(3 * row + c) * ldyk + col] = acc[c][t];
            else YM[(3 * row + c) * ldym + col] = accm[c][t];
        }
}

// L1-resident variant used inside the eigensolver: one 768-thread CTA per SM sweeps a contiguous chunk of an
// ORDERED row list (order[r]: the Morton order of the preconditioner level, or identity), 23 warps taking
// consecutive list entries through a shared-memory ticket, so the X rows gathered by the rows in flight overlap
// and are served by the SM's L1 (nothing is permuted in memory: only the order in which rows are visited changes).
// K values, column ids and M scalars are read once: loaded with L1::no_allocate and pulled into L2 ahead of
// the consumers by the 24th warp (bulk L2 prefetches, one row per lane).
constexpr int SD2_THREADS = 768;
constexpr int SD2_AHEAD = 96;

__device__ __forceinline__ void st_na_f64(double* p, double v) {
    asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void prefetch_l2_run(const void* p, int64_t bytes) {      // any alignment
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uintptr_t lo = a & ~uintptr_t(15), hi = (a + bytes + 15) & ~uintptr_t(15);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(lo), "r"((uint32_t)(hi - lo)) : "memory");
}

template <int CPL>
__global__ void __launch_bounds__(SD2_THREADS, 1)
k_spmm_dual_v2(const int32_t* __restrict__ brow, const int32_t* __restrict__ bcol, const int32_t* __restrict__ order,
               const int32_t* __restrict__ chunk_row, const double* __restrict__ Kval,
               const double* __restrict__ Mblk, const double* __restrict__ X, int64_t ldx, double* __restrict__ YK,
               int64_t ldyk, double* __restrict__ YM, int64_t ldym) {
    __shared__ int s_ticket;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, half = lane >> 4, cl = lane & 15;
    const int r_lo = chunk_row[blockIdx.x], r_hi = chunk_row[blockIdx.x + 1];
    if (tid == 0) s_ticket = 0;
    __syncthreads();
    if (warp == SD2_THREADS / 32 - 1) {
        // prefetcher: stay SD2_AHEAD list entries ahead of the ticket
        for (int base = r_lo; base < r_hi; base += 32) {
            while (base > r_lo + *(volatile int*)&s_ticket + SD2_AHEAD) __nanosleep(200);
            const int r = base + lane;
            if (r < r_hi) {
                const int row = order ? order[r] : r;
                const int64_t b0 = brow[row];
                const int64_t deg = brow[row + 1] - b0;
                if (deg > 0) {
                    prefetch_l2_run(Kval + 9 * b0, 72 * deg);
                    prefetch_l2_run(bcol + b0, 4 * deg);
                    prefetch_l2_run(Mblk + b0, 8 * deg);
                }
            }
        }
        return;
    }
    for (;;) {
        int rr = 0;
        if (lane == 0) rr = atomicAdd(&s_ticket, 1);
        rr = __shfl_sync(0xffffffffu, rr, 0);
        if (r_lo + rr >= r_hi) break;
        const int64_t row = order ? order[r_lo + rr] : r_lo + rr;
        const int64_t b0 = brow[row];
        const int deg = (int)(brow[row + 1] - b0);
        double acc[3][CPL], accm[3][CPL];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int t = 0; t < CPL; ++t) acc[c][t] = accm[c][t] = 0.0;
        const double* kbase = Kval + 9 * b0;
        const int64_t rs = 3 * (int64_t)deg;
        int jn = half < deg ? __ldg(bcol + b0 + half) : 0;        // column id one block ahead (software pipelining)
        for (int p = half; p < deg; p += 2) {
            const int64_t j = jn;
            if (p + 2 < deg) jn = __ldg(bcol + b0 + p + 2);
            const double* xr = X + 3 * j * ldx + cl;
            double x[3][CPL];
#pragma unroll
            for (int d = 0; d < 3; ++d)
#pragma unroll
                for (int t = 0; t < CPL; ++t) x[d][t] = __ldg(xr + d * ldx + 16 * t);
            const double* kp = kbase + 3 * p;
            double k[3][3];
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int d = 0; d < 3; ++d) k[c][d] = __ldg(kp + c * rs + d);      // line reuse across p: normal L1 policy
                                                                                     // (no_allocate and evict_first both measured slower)
            const double m = __ldg(Mblk + b0 + p);
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int t = 0; t < CPL; ++t) {
                    double a = acc[c][t];
                    a = fma(k[c][0], x[0][t], a);
                    a = fma(k[c][1], x[1][t], a);
                    a = fma(k[c][2], x[2][t], a);
                    acc[c][t] = a;
                    accm[c][t] = fma(m, x[c][t], accm[c][t]);
                }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int t = 0; t < CPL; ++t) {
                acc[c][t] += __shfl_xor_sync(0xffffffffu, acc[c][t], 16);
                accm[c][t] += __shfl_xor_sync(0xffffffffu, accm[c][t], 16);
            }
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int t = 0; t < CPL; ++t) {
                const int64_t col = cl + 16 * t;
                if (half == 0) st_na_f64(YK + (3 * row + c) * ldyk + col, acc[c][t]);
                else st_na_f64(YM + (3 * row + c) * ldym + col, accm[c][t]);
            }
    }
}

// K Z and M Z for a block Z that is exactly representable in FP32 -- the output of the FP32 preconditioner, which
// is what LOBPCG's new search directions W are (csrc/lobpcg.cu).  The gathered operand is the level's own fp32 block
// in ITS (Morton) numbering: rows r' of Z, column ids bcolP (already mapped to that numbering, in the block order of
// the matrix row perm[r']) -- 4 bytes per gathered value instead of 8, Morton-contiguous rows instead of rows strided
// through the n x 3m iterate buffer.  Matrix values stay FP64, the accumulation is FP64, results go to rows
// 3 perm[r'] + c of YK / YM.
// Mapping (the one of k_spmm32v, csrc/precond32.cu): one CTA per SM sweeps a contiguous chunk of the Morton row list,
// warps take rows through a shared-memory ticket; LPR lanes cover the C = LPR * CPT columns with 128/64-bit loads, the
// 32 / LPR lane groups walk alternate blocks of the row, two blocks per group in flight (the FP64 kernel above keeps
// one block per half-warp in flight and is bound by load latency: long-scoreboard stalls at 37 % occupancy).
constexpr int SZ_THREADS = 384;   // 168 registers; 512 threads (128 registers, no spills) measured the same, 640 (96, spills) 20 % slower
constexpr int SZ_PF = 48;        // rows of look-ahead of the L2 prefetch

template <int LPR, int CPT>
__device__ __forceinline__ void ld_row_f32(const float* __restrict__ row, int l, float* out) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(row) + l);
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
    if constexpr (CPT == 6) {
        const float2 w = __ldg(reinterpret_cast<const float2*>(row + 4 * LPR) + l);
        out[4] = w.x; out[5] = w.y;
    }
}
template <int LPR, int CPT>
__device__ __forceinline__ void st_row_f64(double* row, int l, const double* v) {
    double2* p = reinterpret_cast<double2*>(row + 4 * l);
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v[0]), "d"(v[1]) : "memory");
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p + 1), "d"(v[2]), "d"(v[3]) : "memory");
    if constexpr (CPT == 6)
        asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(row + 4 * LPR + 2 * l), "d"(v[4]), "d"(v[5]) : "memory");
}

template <int LPR, int CPT>
__global__ void __launch_bounds__(SZ_THREADS, 1)
k_spmm_dual_z32(const int32_t* __restrict__ brow, const int32_t* __restrict__ browP, const int32_t* __restrict__ bcolP,
                const int32_t* __restrict__ perm, const int32_t* __restrict__ chunk_row,
                const double* __restrict__ Kval, const double* __restrict__ Mblk, const float* __restrict__ Z,
                double* __restrict__ YK, int64_t ldyk, double* __restrict__ YM, int64_t ldym) {
    constexpr int C = LPR * CPT, NG = 32 / LPR;
    __shared__ int s_ticket;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane / LPR, l = lane % LPR;
    const int r_lo = chunk_row[blockIdx.x], r_hi = chunk_row[blockIdx.x + 1];
    if (tid == 0) s_ticket = 0;
    auto prefetch_row = [&](int r) {
        const int row = perm ? perm[r] : r;
        const int64_t b0 = brow[row];
        const int64_t deg = brow[row + 1] - b0;
        if (deg > 0) {
            prefetch_l2_run(Kval + 9 * b0, 72 * deg);
            prefetch_l2_run(bcolP + browP[r], 4 * deg);
            prefetch_l2_run(Mblk + b0, 8 * deg);
        }
    };
    if (lane == 0) {
        for (int q = warp; q < SZ_PF && r_lo + q < r_hi; q += SZ_THREADS / 32) prefetch_row(r_lo + q);
    }
    __syncthreads();
    for (;;) {
        int rr = 0;
        if (lane == 0) rr = atomicAdd(&s_ticket, 1);
        rr = __shfl_sync(0xffffffffu, rr, 0);
        const int rp = r_lo + rr;
        if (rp >= r_hi) break;
        if (lane == 0 && rp + SZ_PF < r_hi) prefetch_row(rp + SZ_PF);
        const int64_t row = perm ? perm[rp] : rp;
        const int64_t b0 = brow[row];
        const int deg = (int)(brow[row + 1] - b0);
        const int32_t* __restrict__ cp = bcolP + browP[rp];
        const double* __restrict__ kb = Kval + 9 * b0;
        const double* __restrict__ mb = Mblk + b0;
        const int64_t rs = 3 * (int64_t)deg;
        double acc[3][CPT], accm[3][CPT];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int t = 0; t < CPT; ++t) acc[c][t] = accm[c][t] = 0.0;
        auto fma_block = [&](const float (&x)[3][CPT], const double (&k)[9], double mv) {
#pragma unroll
            for (int t = 0; t < CPT; ++t) {
                const double x0 = (double)x[0][t], x1 = (double)x[1][t], x2 = (double)x[2][t];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    double a = acc[c][t];
                    a = fma(k[3 * c], x0, a);
                    a = fma(k[3 * c + 1], x1, a);
                    a = fma(k[3 * c + 2], x2, a);
                    acc[c][t] = a;
                }
                accm[0][t] = fma(mv, x0, accm[0][t]);
                accm[1][t] = fma(mv, x1, accm[1][t]);
                accm[2][t] = fma(mv, x2, accm[2][t]);
            }
        };
        int p = g;
        int ja = p < deg ? __ldg(cp + p) : 0;
        int jb = p + NG < deg ? __ldg(cp + p + NG) : 0;
        for (; p + NG < deg; p += 2 * NG) {
            const int jan = p + 2 * NG < deg ? __ldg(cp + p + 2 * NG) : 0;
            const int jbn = p + 3 * NG < deg ? __ldg(cp + p + 3 * NG) : 0;
            const float* xa = Z + (int64_t)3 * ja * C;
            const float* xb = Z + (int64_t)3 * jb * C;
            float x[3][CPT], y[3][CPT];
            double ka[9], kq[9];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                ld_row_f32<LPR, CPT>(xa + d * C, l, x[d]);
                ld_row_f32<LPR, CPT>(xb + d * C, l, y[d]);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    ka[3 * c + d] = __ldg(kb + c * rs + 3 * p + d);
                    kq[3 * c + d] = __ldg(kb + c * rs + 3 * (p + NG) + d);
                }
            const double ma = __ldg(mb + p), mq = __ldg(mb + p + NG);
            fma_block(x, ka, ma);
            fma_block(y, kq, mq);
            ja = jan;
            jb = jbn;
        }
        if (p < deg) {
            const float* xa = Z + (int64_t)3 * ja * C;
            float x[3][CPT];
            double ka[9];
#pragma unroll
            for (int d = 0; d < 3; ++d) ld_row_f32<LPR, CPT>(xa + d * C, l, x[d]);
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int d = 0; d < 3; ++d) ka[3 * c + d] = __ldg(kb + c * rs + 3 * p + d);
            fma_block(x, ka, __ldg(mb + p));
        }
#pragma unroll
        for (int off = LPR; off < 32; off <<= 1)
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int t = 0; t < CPT; ++t) {
                    acc[c][t] += __shfl_xor_sync(0xffffffffu, acc[c][t], off);
                    accm[c][t] += __shfl_xor_sync(0xffffffffu, accm[c][t], off);
                }
        // lane group c < 3 stores component row c of K Z, group (c + 1) % NG ... of M Z (NG >= 4), spreading the stores
        if (g < 3) {
            double v[CPT];
#pragma unroll
            for (int t = 0; t < CPT; ++t) v[t] = g == 0 ? acc[0][t] : (g == 1 ? acc[1][t] : acc[2][t]);
            st_row_f64<LPR, CPT>(YK + (3 * row + g) * ldyk, l, v);
        }
        const int gm = NG - 1 - g;
        if (gm < 3) {
            double v[CPT];
#pragma unroll
            for (int t = 0; t < CPT; ++t) v[t] = gm == 0 ? accm[0][t] : (gm == 1 ? accm[1][t] : accm[2][t]);
            st_row_f64<LPR, CPT>(YM + (3 * row + gm) * ldym, l, v);
        }
    }
}

static inline unsigned row_blocks(int64_t n_nodes) { return (unsigned)ceil_div(n_nodes * 32, 256); }

#define DS_DISPATCH_CPL(cpl, ...)                                              \
    switch (cpl) {                                                             \
        case 1: { constexpr int CPL = 1; __VA_ARGS__; } break;                 \
        case 2: { constexpr int CPL = 2; __VA_ARGS__; } break;                 \
        case 3: { constexpr int CPL = 3; __VA_ARGS__; } break;                 \
        case 4: { constexpr int CPL = 4; __VA_ARGS__; } break;                 \
        case 5: { constexpr int CPL = 5; __VA_ARGS__; } break;                 \
        case 6: { constexpr int CPL = 6; __VA_ARGS__; } break;                 \
        case 7: { constexpr int CPL = 7; __VA_ARGS__; } break;                 \
        case 8: { constexpr int CPL = 8; __VA_ARGS__; } break;                 \
        default: set_error("ncols must be a multiple of 16 in [16,128]"); return DS_ERR_ARG; \
    }

int spmm_km(const int32_t* brow, const int32_t* bcol, int64_t n_nodes, const double* Kval, const double* Mblk,
            double shift, const double* X, int64_t ldx, int ncols, double alpha, double beta, const double* Y0,
            int64_t ldy0, double* Y, int64_t ldy, cudaStream_t stream) {
    DS_REQUIRE(ncols > 0 && ncols % 16 == 0 && ncols <= 128, "spmm: ncols=%d must be a multiple of 16 <= 128", ncols);
    DS_REQUIRE(brow && bcol && X && Y, "spmm: null argument");
    DS_REQUIRE(Kval || Mblk, "spmm: both Kval and Mblk are NULL");
    DS_REQUIRE(beta == 0.0 || Y0, "spmm: beta != 0 needs Y0");
    if (beta == 0.0) { Y0 = Y; ldy0 = ldy; }
    unsigned g = row_blocks(n_nodes);
    int cpl = ncols / 16;
    ProfScope prof(PROF_SPMM, stream);
    if (Kval && Mblk) {
        DS_DISPATCH_CPL(cpl, (k_spmm<CPL, true, true><<<g, 256, 0, stream>>>(brow, bcol, n_nodes, Kval, Mblk, shift, X,
                                                                            ldx, alpha, beta, Y0, ldy0, Y, ldy)));
    } else if (Kval) {
        DS_DISPATCH_CPL(cpl, (k_spmm<CPL, true, false><<<g, 256, 0, stream>>>(brow, bcol, n_nodes, Kval, Mblk, shift, X,
                                                                             ldx, alpha, beta, Y0, ldy0, Y, ldy)));
    } else {
        // pure M product: shift scales it
        DS_DISPATCH_CPL(cpl, (k_spmm<CPL, false, true><<<g, 256, 0, stream>>>(brow, bcol, n_nodes, Kval, Mblk, shift, X,
                                                                             ldx, alpha, beta, Y0, ldy0, Y, ldy)));
    }
    DS_LAUNCH_CHECK();
    return DS_OK;
}

int spmm_dual(const int32_t* brow, const int32_t* bcol, int64_t n_nodes, const double* Kval, const double* Mblk,
              const double* X, int64_t ldx, int ncols, double* YK, int64_t ldyk, double* YM, int64_t ldym,
              cudaStream_t stream, const int32_t* order, const int32_t* chunk_row, int nchunks) {
    DS_REQUIRE(ncols > 0 && ncols % 16 == 0 && ncols <= 128, "spmm: ncols=%d must be a multiple of 16 <= 128", ncols);
    DS_REQUIRE(brow && bcol && Kval && Mblk && X && YK && YM, "spmm_k_and_m: null argument");
    unsigned g = row_blocks(n_nodes);
    int cpl = ncols / 16;
    ProfScope prof(PROF_SPMM, stream);
    {
        int32_t nb = 0;      // nnzb is not an argument of this entry point: accounted only while profiling (one 4-byte D2H)
        if (prof_enabled(PROF_SPMM) && cudaMemcpyAsync(&nb, brow + n_nodes, 4, cudaMemcpyDeviceToHost, stream) == cudaSuccess &&
            cudaStreamSynchronize(stream) == cudaSuccess)
            prof_account(PROF_SPMM, (double)nb * 84.0 + (double)n_nodes * 4.0 + 3.0 * n_nodes * ncols * 24.0, 2.0 * (double)nb * 12.0 * ncols);
    }
    if (chunk_row && nchunks > 0 && cpl <= 3) {
        // ordered sweep, one CTA per chunk (= per SM), all of the SM's 256 KB as L1
        auto go = [&](auto kern) -> int {
            static bool carved = false;
            if (!carved) {
                DS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1));
                carved = true;
            }
            kern<<<nchunks, SD2_THREADS, 0, stream>>>(brow, bcol, order, chunk_row, Kval, Mblk, X, ldx, YK, ldyk, YM, ldym);
            DS_LAUNCH_CHECK();
            return DS_OK;
        };
        return cpl == 1 ? go(k_spmm_dual_v2<1>) : (cpl == 2 ? go(k_spmm_dual_v2<2>) : go(k_spmm_dual_v2<3>));
    }
    DS_DISPATCH_CPL(cpl, (k_spmm_dual<CPL><<<g, 256, 0, stream>>>(brow, bcol, n_nodes, Kval, Mblk, X, ldx, YK, ldyk, YM,
                                                                  ldym)));
    DS_LAUNCH_CHECK();
    return DS_OK;
}

int spmm_dual_z32(const int32_t* brow, const int32_t* browP, const int32_t* bcolP, const int32_t* perm,
                  const int32_t* chunk_row, int nchunks, int64_t n_nodes, const double* Kval, const double* Mblk,
                  const float* Z, int ncols, double* YK, int64_t ldyk, double* YM, int64_t ldym, cudaStream_t stream,
                  int64_t nnzb) {
    DS_REQUIRE(ncols > 0 && ncols % 16 == 0 && ncols <= 48, "spmm_dual_z32: ncols=%d must be 16, 32 or 48", ncols);
    DS_REQUIRE(brow && browP && bcolP && chunk_row && nchunks > 0 && Kval && Mblk && Z && YK && YM, "spmm_dual_z32: null argument");
    ProfScope prof(PROF_SPMM, stream);
    // algorithmic bytes: 9 fp64 K values + 1 fp64 M scalar + 4 B column id per block, row pointers, Z once (fp32), K Z and M Z out
    prof_account(PROF_SPMM, (double)nnzb * 84.0 + (double)n_nodes * 12.0 + 3.0 * n_nodes * ncols * (4.0 + 16.0),
                 2.0 * (double)nnzb * 12.0 * ncols);
    auto go = [&](auto kern) -> int {
        static bool carved = false;
        if (!carved) {
            DS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1));
            carved = true;
        }
        kern<<<nchunks, SZ_THREADS, 0, stream>>>(brow, browP, bcolP, perm, chunk_row, Kval, Mblk, Z, YK, ldyk, YM, ldym);
        DS_LAUNCH_CHECK();
        return DS_OK;
    };
    DS_REQUIRE(ldyk % 2 == 0 && ldym % 2 == 0 && (uintptr_t)YK % 16 == 0 && (uintptr_t)YM % 16 == 0 && (uintptr_t)Z % 16 == 0,
               "spmm_dual_z32: outputs and Z must be 16-byte aligned with even leading dimensions");
    return ncols == 16 ? go(k_spmm_dual_z32<4, 4>) : (ncols == 32 ? go(k_spmm_dual_z32<8, 4>) : go(k_spmm_dual_z32<8, 6>));
}

}  // namespace ds

using namespace ds;

extern "C" int ds_spmm_km(const int32_t* brow, const int32_t* bcol, int64_t n_nodes, const double* Kval,
                          const double* Mblk, double shift, const double* X, int64_t ldx, int ncols, double alpha,
                          double beta, const double* Y0, int64_t ldy0, double* Y, int64_t ldy, void* stream) {
    return spmm_km(brow, bcol, n_nodes, Kval, Mblk, shift, X, ldx, ncols, alpha, beta, Y0, ldy0, Y, ldy,
                   (cudaStream_t)stream);
}

extern "C" int ds_spmm_k_and_m(const int32_t* brow, const int32_t* bcol, int64_t n_nodes, const double* Kval,
                               const double* Mblk, const double* X, int64_t ldx, int ncols, double* YK, int64_t ldyk,
                               double* YM, int64_t ldym, void* stream) {
    return spmm_dual(brow, bcol, n_nodes, Kval, Mblk, X, ldx, ncols, YK, ldyk, YM, ldym, (cudaStream_t)stream);
}
