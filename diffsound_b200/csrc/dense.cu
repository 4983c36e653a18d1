// Dense tall-skinny kernels of the Rayleigh-Ritz step, FP64.
//
// Reference behaviour replaced (under /root/reference/src/lobpcg): the torch.matmul /
// qform Gram products S^T (A S) (_linalg_utils.py:63-73, _lobpcg.py:460,479-525), the
// basis updates X = S Z (_lobpcg.py:463-466) and torch.linalg.cholesky/eigh on the
// small projected problem (_lobpcg.py:507-525, _linalg_utils.py:87-96).
//
//  * gram_f64:   G = A^T B.  Row tiles are streamed into shared memory by the TMA
//                unit (1-D cp.async.bulk per row, mbarrier completion, 3 stages);
//                the product runs on the FP64 tensor pipe (mma.sync m8n8k4 f64 ->
//                SASS DMMA); per-CTA partials are reduced in a fixed order
//                (deterministic, no atomics).
//  * block_gemm_f64:  Y = beta Y + A C with C (p x q) resident in shared memory,
//                A fragments read straight from global in DMMA layout (one 32 B
//                sector per lane quad).
//  (the small generalised eigen-solve of the Rayleigh-Ritz step lives in eigh.cu)
#include "common.cuh"
#include "../../include/diffsound_sm100.h"
#include "kernels.cuh"
#include "ptx.cuh"
#include <algorithm>

namespace ds {

// ---------------------------------------------------------------------------
// Gram: G[p x q] = A^T B
// ---------------------------------------------------------------------------
constexpr int GRAM_ROWS = 32;      // rows per stage
constexpr int GRAM_STAGES = 3;
constexpr int GRAM_THREADS = 256;  // 8 warps
constexpr int GRAM_MAX_CTAS = 296; // 2 per SM

__host__ __device__ inline int pad8mod16(int p) {  // smallest s >= p with s % 16 == 8
    int s = (p / 16) * 16 + 8;
    return s >= p ? s : s + 16;
}

template <int MAXT>
__global__ void __launch_bounds__(GRAM_THREADS)
k_gram(const double* __restrict__ A, int64_t lda, int p, const double* __restrict__ B, int64_t ldb, int q,
       int64_t n, double* __restrict__ partial) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int ps = pad8mod16(p), qs = pad8mod16(q);
    double* As = reinterpret_cast<double*>(smem_raw);                       // [STAGES][ROWS][ps]
    double* Bs = As + GRAM_STAGES * GRAM_ROWS * ps;                          // [STAGES][ROWS][qs]
    uint64_t* full = reinterpret_cast<uint64_t*>(Bs + GRAM_STAGES * GRAM_ROWS * qs);
    uint64_t* empty = full + GRAM_STAGES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_tiles = (n + GRAM_ROWS - 1) / GRAM_ROWS;
    // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const int64_t my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (tid == 0) {
        for (int s = 0; s < GRAM_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], GRAM_THREADS / 32);
        }
        fence_barrier_init();
    }
    __syncthreads();

    const int tp = p >> 3, tq = q >> 3, ntile = tp * tq;
    double acc[MAXT][2];
#pragma unroll
    for (int t = 0; t < MAXT; ++t) acc[t][0] = acc[t][1] = 0.0;

    auto issue = [&](int64_t it) {  // executed by warp 0, all lanes
        int s = (int)(it % GRAM_STAGES);
        int64_t tile = blockIdx.x + it * (int64_t)gridDim.x;
        int64_t r0 = tile * GRAM_ROWS;
        int rows = (int)min((int64_t)GRAM_ROWS, n - r0);
        if (it >= GRAM_STAGES) mbar_wait(&empty[s], (uint32_t)(((it / GRAM_STAGES) - 1) & 1));
        if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)(rows * (p + q) * sizeof(double)));
        __syncwarp();
        if (lane < rows) {
            tma_load_1d(As + ((size_t)s * GRAM_ROWS + lane) * ps, A + (r0 + lane) * lda, p * 8, &full[s]);
            tma_load_1d(Bs + ((size_t)s * GRAM_ROWS + lane) * qs, B + (r0 + lane) * ldb, q * 8, &full[s]);
        }
    };

    if (warp == 0) {
        for (int64_t it = 0; it < min((int64_t)GRAM_STAGES, my_tiles); ++it) issue(it);
    }
    for (int64_t it = 0; it < my_tiles; ++it) {
        int s = (int)(it % GRAM_STAGES);
        mbar_wait(&full[s], (uint32_t)((it / GRAM_STAGES) & 1));
        int64_t tile = blockIdx.x + it * (int64_t)gridDim.x;
        int rows = (int)min((int64_t)GRAM_ROWS, n - tile * GRAM_ROWS);
        const double* as = As + (size_t)s * GRAM_ROWS * ps;
        const double* bs = Bs + (size_t)s * GRAM_ROWS * qs;
        const int kk = lane & 3, mm = lane >> 2;
#pragma unroll
        for (int t = 0; t < MAXT; ++t) {
            int idx = warp + 8 * t;
            if (idx < ntile) {
                int ti = idx / tq, tj = idx - ti * tq;
                const double* ap = as + kk * ps + 8 * ti + mm;
                const double* bp = bs + kk * qs + 8 * tj + mm;
#pragma unroll
                for (int k0 = 0; k0 < GRAM_ROWS; k0 += 4) {
                    bool ok = (k0 + kk) < rows;
                    double a = ok ? ap[k0 * ps] : 0.0;
                    double b = ok ? bp[k0 * qs] : 0.0;
                    dmma_m8n8k4(acc[t][0], acc[t][1], a, b);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (warp == 0 && it + GRAM_STAGES < my_tiles) issue(it + GRAM_STAGES);
    }
    // write partial tile sums: partial[blockIdx.x][p][q]
    double* out = partial + (size_t)blockIdx.x * p * q;
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
        int idx = warp + 8 * t;
        if (idx < ntile) {
            int ti = idx / tq, tj = idx - ti * tq;
            int r = 8 * ti + (lane >> 2), c = 8 * tj + 2 * (lane & 3);
            out[r * q + c] = acc[t][0];
            out[r * q + c + 1] = acc[t][1];
        }
    }
}

// one warp per output element: lane l sums partials l, l + 32, ... and the lanes are combined by a fixed butterfly
// (deterministic); the one-thread-per-element version walked its 296 partials alone and took 77 us on 9 CTAs
__global__ void __launch_bounds__(256)
k_gram_reduce(const double* __restrict__ partial, int nparts, int p, int q, double* __restrict__ G, int64_t ldg) {
    const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (idx >= p * q) return;
    double s = 0.0;
    for (int c = lane; c < nparts; c += 32) s += partial[(size_t)c * p * q + idx];
    s = warp_sum(s);
    if (lane == 0) G[(int64_t)(idx / q) * ldg + (idx % q)] = s;
}

static int gram_ctas(int64_t n) {
    int64_t tiles = (n + GRAM_ROWS - 1) / GRAM_ROWS;
    return (int)(tiles < GRAM_MAX_CTAS ? tiles : GRAM_MAX_CTAS);
}

int64_t gram_scratch_elems(int p, int q) { return (int64_t)GRAM_MAX_CTAS * p * q; }

int gram_f64(const double* A, int64_t lda, int p, const double* B, int64_t ldb, int q, int64_t n, double* G,
             int64_t ldg, double* partial, cudaStream_t stream) {
    DS_REQUIRE(p > 0 && q > 0 && p % 8 == 0 && q % 8 == 0 && p <= 64 && q <= 64,
               "gram: p=%d, q=%d must be multiples of 8 in [8,64]", p, q);
    DS_REQUIRE(A && B && G && partial, "gram: null argument");
    DS_REQUIRE(lda % 2 == 0 && ldb % 2 == 0 && ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0),
               "gram: operands must be 16-byte aligned with even leading dimensions");
    DS_REQUIRE(n > 0, "gram: n must be positive");
    int ps = pad8mod16(p), qs = pad8mod16(q);
    size_t smem = (size_t)GRAM_STAGES * GRAM_ROWS * (ps + qs) * sizeof(double) + 2 * GRAM_STAGES * sizeof(uint64_t);
    int ctas = gram_ctas(n);
    ProfScope prof(PROF_GRAM, stream);
    prof_account(PROF_GRAM, (double)n * (p + q) * 8.0, 2.0 * (double)n * p * q);
    int ntile = (p / 8) * (q / 8);
    int maxt = (ntile + 7) / 8;
    auto launch = [&](auto kern) -> int {
        DS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<ctas, GRAM_THREADS, smem, stream>>>(A, lda, p, B, ldb, q, n, partial);
        DS_LAUNCH_CHECK();
        return DS_OK;
    };
    if (maxt <= 2) DS_TRY(launch(k_gram<2>));
    else if (maxt <= 5) DS_TRY(launch(k_gram<5>));
    else DS_TRY(launch(k_gram<8>));
    k_gram_reduce<<<(p * q * 32 + 255) / 256, 256, 0, stream>>>(partial, ctas, p, q, G, ldg);
    DS_LAUNCH_CHECK();
    return DS_OK;
}

// ---------------------------------------------------------------------------
// Block GEMM: Y (n x q) = beta Y + A (n x p) C (p x q)
// ---------------------------------------------------------------------------
constexpr int BG_THREADS = 256;  // 8 warps, 8 rows each -> 64 rows per CTA pass

template <int QT>  // q/8 tiles per warp
__global__ void __launch_bounds__(BG_THREADS)
k_block_gemm(const double* __restrict__ A, int64_t lda, int p, const double* __restrict__ C, int64_t ldc, int q,
             int64_t n, double alpha, double beta, double* __restrict__ Y, int64_t ldy) {
    extern __shared__ __align__(16) double Cs[];  // [p][qs]
    const int qs = pad8mod16(q);
    for (int t = threadIdx.x; t < p * q; t += blockDim.x) {
        int r = t / q, c = t - r * q;
        Cs[r * qs + c] = alpha * C[(int64_t)r * ldc + c];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kk = lane & 3, mm = lane >> 2;
    const int64_t n_strips = (n + 7) / 8;
    for (int64_t strip = blockIdx.x * 8 + warp; strip < n_strips; strip += (int64_t)gridDim.x * 8) {
        int64_t row = strip * 8 + mm;
        bool ok = row < n;
        const double* ap = A + (ok ? row : 0) * lda + kk;
        double acc[QT][2];
#pragma unroll
        for (int t = 0; t < QT; ++t) acc[t][0] = acc[t][1] = 0.0;
        for (int k0 = 0; k0 < p; k0 += 16) {
            double a[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u] = (ok && (k0 + 4 * u) < p) ? __ldg(ap + k0 + 4 * u) : 0.0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (k0 + 4 * u < p) {
                    const double* bp = Cs + (k0 + 4 * u + kk) * qs + mm;
#pragma unroll
                    for (int t = 0; t < QT; ++t) dmma_m8n8k4(acc[t][0], acc[t][1], a[u], bp[8 * t]);
                }
            }
        }
        if (ok) {
            double* yp = Y + row * ldy + 2 * kk;
#pragma unroll
            for (int t = 0; t < QT; ++t) {
                double2 v = make_double2(acc[t][0], acc[t][1]);
                if (beta != 0.0) {
                    double2 o = *reinterpret_cast<const double2*>(yp + 8 * t);
                    v.x = fma(beta, o.x, v.x);
                    v.y = fma(beta, o.y, v.y);
                }
                *reinterpret_cast<double2*>(yp + 8 * t) = v;
            }
        }
    }
}

int block_gemm_f64(const double* A, int64_t lda, int p, const double* C, int64_t ldc, int q, int64_t n, double alpha,
                   double beta, double* Y, int64_t ldy, cudaStream_t stream) {
    DS_REQUIRE(p > 0 && q > 0 && p % 4 == 0 && q % 8 == 0 && q <= 64 && p <= 192,
               "block_gemm: p=%d (mult of 4, <=192), q=%d (mult of 8, <=64)", p, q);
    DS_REQUIRE(A && C && Y, "block_gemm: null argument");
    DS_REQUIRE(ldy % 2 == 0 && ((uintptr_t)Y % 16 == 0), "block_gemm: Y must be 16-byte aligned, even ldy");
    DS_REQUIRE(A != Y, "block_gemm: A must not alias Y");
    int qs = pad8mod16(q);
    ProfScope prof(PROF_GEMM, stream);
    prof_account(PROF_GEMM, (double)n * (p + (beta != 0.0 ? 2 : 1) * q) * 8.0, 2.0 * (double)n * p * q);
    size_t smem = (size_t)p * qs * sizeof(double);
    int64_t strips = (n + 7) / 8;
    int ctas = (int)std::min<int64_t>((strips + 7) / 8, 148 * 4);
    auto launch = [&](auto kern) -> int {
        DS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<ctas, BG_THREADS, smem, stream>>>(A, lda, p, C, ldc, q, n, alpha, beta, Y, ldy);
        DS_LAUNCH_CHECK();
        return DS_OK;
    };
    switch (q / 8) {
        case 1: DS_TRY(launch(k_block_gemm<1>)); break;
        case 2: DS_TRY(launch(k_block_gemm<2>)); break;
        case 3: DS_TRY(launch(k_block_gemm<3>)); break;
        case 4: DS_TRY(launch(k_block_gemm<4>)); break;
        case 5: DS_TRY(launch(k_block_gemm<5>)); break;
        case 6: DS_TRY(launch(k_block_gemm<6>)); break;
        case 7: DS_TRY(launch(k_block_gemm<7>)); break;
        default: DS_TRY(launch(k_block_gemm<8>)); break;
    }
    return DS_OK;
}

// ---------------------------------------------------------------------------
// Fused Rayleigh-Ritz update of one LOBPCG step, all three wide buffers in one launch:
//   Ynew[:, 0:m]           = A[:, 0:prow] C1      (the new X block)
//   Ynew[:, 2m:2m + q2]    = A[:, m:prow] C2      (the new P block, active columns only)
// for A in {S, KS, MS} (blockIdx.y).  A is read ONCE for both products (the separate GEMMs read the
// W and P columns twice); both coefficient matrices sit in shared memory; the A fragments of the next
// 16 columns are in flight while the DMMAs of the current 16 run.
// ---------------------------------------------------------------------------
struct RRUpdateArgs {
    const double* A[3];
    double* Y[3];
    int64_t lda, ldy, n;
    const double* C1;       // prow x m
    const double* C2;       // (prow - m) x q2
    int64_t ldc;
    int prow, m, q2;
};

template <int QT1, int QT2>
__global__ void __launch_bounds__(BG_THREADS)
k_rr_update(const __grid_constant__ RRUpdateArgs g) {
    extern __shared__ __align__(16) double Cs[];  // C1s [prow][qs1] | C2s [prow - m][qs2]
    constexpr int q1 = 8 * QT1, q2 = 8 * QT2;
    constexpr int NS = 2;                          // row strips per warp: every B fragment read from shared memory
                                                   // feeds NS DMMAs (the one-strip version ran at 96 % of the
                                                   // shared-memory pipe and 74 % of the FP64 tensor pipe)
    const int qs1 = pad8mod16(q1), qs2 = QT2 ? pad8mod16(q2) : 0;
    double* C1s = Cs;
    double* C2s = Cs + (size_t)g.prow * qs1;
    for (int t = threadIdx.x; t < g.prow * q1; t += blockDim.x) {
        const int r = t / q1, c = t - r * q1;
        C1s[r * qs1 + c] = g.C1[(int64_t)r * g.ldc + c];
    }
    if (QT2)
        for (int t = threadIdx.x; t < (g.prow - g.m) * q2; t += blockDim.x) {
            const int r = t / q2, c = t - r * q2;
            C2s[r * qs2 + c] = g.C2[(int64_t)r * g.ldc + c];
        }
    __syncthreads();
    const double* __restrict__ A = g.A[blockIdx.y];
    double* __restrict__ Y = g.Y[blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kk = lane & 3, mm = lane >> 2;
    const int64_t n_strips = (g.n + 7) / 8;
    const int p = g.prow, m = g.m;
    for (int64_t strip = ((int64_t)blockIdx.x * 8 + warp) * NS; strip < n_strips; strip += (int64_t)gridDim.x * 8 * NS) {
        bool ok[NS];
        const double* ap[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int64_t row = (strip + s) * 8 + mm;
            ok[s] = row < g.n;
            ap[s] = A + (ok[s] ? row : 0) * g.lda + kk;
        }
        double acc1[NS][QT1][2], acc2[NS][QT2 ? QT2 : 1][2];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
#pragma unroll
            for (int t = 0; t < QT1; ++t) acc1[s][t][0] = acc1[s][t][1] = 0.0;
#pragma unroll
            for (int t = 0; t < (QT2 ? QT2 : 1); ++t) acc2[s][t][0] = acc2[s][t][1] = 0.0;
        }
        double a[NS][4], an[NS][4];
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int u = 0; u < 4; ++u) a[s][u] = (ok[s] && 4 * u < p) ? __ldg(ap[s] + 4 * u) : 0.0;
        for (int k0 = 0; k0 < p; k0 += 16) {
#pragma unroll
            for (int s = 0; s < NS; ++s)
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    an[s][u] = (ok[s] && (k0 + 16 + 4 * u) < p) ? __ldg(ap[s] + k0 + 16 + 4 * u) : 0.0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = k0 + 4 * u;
                if (k < p) {
                    const double* bp = C1s + (k + kk) * qs1 + mm;
#pragma unroll
                    for (int t = 0; t < QT1; ++t) {
                        const double b = bp[8 * t];
#pragma unroll
                        for (int s = 0; s < NS; ++s) dmma_m8n8k4(acc1[s][t][0], acc1[s][t][1], a[s][u], b);
                    }
                    if (QT2 && k >= m) {
                        const double* cp = C2s + (k - m + kk) * qs2 + mm;
#pragma unroll
                        for (int t = 0; t < QT2; ++t) {
                            const double b = cp[8 * t];
#pragma unroll
                            for (int s = 0; s < NS; ++s) dmma_m8n8k4(acc2[s][t][0], acc2[s][t][1], a[s][u], b);
                        }
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < NS; ++s)
#pragma unroll
                for (int u = 0; u < 4; ++u) a[s][u] = an[s][u];
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            if (!ok[s]) continue;
            double* yp = Y + ((strip + s) * 8 + mm) * g.ldy + 2 * kk;
#pragma unroll
            for (int t = 0; t < QT1; ++t)
                *reinterpret_cast<double2*>(yp + 8 * t) = make_double2(acc1[s][t][0], acc1[s][t][1]);
            if (QT2) {
                double* zp = yp + 2 * m;
#pragma unroll
                for (int t = 0; t < QT2; ++t)
                    *reinterpret_cast<double2*>(zp + 8 * t) = make_double2(acc2[s][t][0], acc2[s][t][1]);
            }
        }
    }
}

int rr_update_f64(const double* const A[3], int64_t lda, int prow, int m, const double* C1, const double* C2, int q2,
                  int64_t ldc, int64_t n, double* const Y[3], int64_t ldy, cudaStream_t stream) {
    DS_REQUIRE(m > 0 && m % 16 == 0 && m <= 48 && prow >= m && prow % 4 == 0 && prow <= 144,
               "rr_update: m=%d (16, 32 or 48), prow=%d (multiple of 4 in [m, 144])", m, prow);
    DS_REQUIRE(q2 >= 0 && q2 % 16 == 0 && q2 <= 48 && (q2 == 0 || prow > m), "rr_update: q2=%d must be 0, 16, 32 or 48", q2);
    DS_REQUIRE(C1 && (q2 == 0 || C2), "rr_update: null coefficient matrix");
    RRUpdateArgs g;
    for (int b = 0; b < 3; ++b) {
        DS_REQUIRE(A[b] && Y[b] && A[b] != Y[b], "rr_update: bad buffer %d", b);
        DS_REQUIRE((uintptr_t)Y[b] % 16 == 0, "rr_update: Y must be 16-byte aligned");
        g.A[b] = A[b];
        g.Y[b] = Y[b];
    }
    DS_REQUIRE(ldy % 2 == 0, "rr_update: ldy must be even");
    g.lda = lda; g.ldy = ldy; g.n = n; g.C1 = C1; g.C2 = C2; g.ldc = ldc; g.prow = prow; g.m = m; g.q2 = q2;
    ProfScope prof(PROF_GEMM, stream);
    const size_t smem = ((size_t)prow * pad8mod16(m) + (q2 ? (size_t)(prow - m) * pad8mod16(q2) : 0)) * sizeof(double);
    const int64_t strips = (n + 7) / 8;
    const int per_buf = (int)std::min<int64_t>((strips + 15) / 16, 148 * 2);
    auto launch = [&](auto kern) -> int {
        DS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<dim3(per_buf, 3), BG_THREADS, smem, stream>>>(g);
        DS_LAUNCH_CHECK();
        return DS_OK;
    };
#define DS_RR_CASE(Q1, Q2) if (m == 8 * Q1 && q2 == 8 * Q2) return launch(k_rr_update<Q1, Q2>)
    DS_RR_CASE(2, 0); DS_RR_CASE(2, 2);
    DS_RR_CASE(4, 0); DS_RR_CASE(4, 2); DS_RR_CASE(4, 4);
    DS_RR_CASE(6, 0); DS_RR_CASE(6, 2); DS_RR_CASE(6, 4); DS_RR_CASE(6, 6);
#undef DS_RR_CASE
    set_error("rr_update: unsupported shape m=%d q2=%d", m, q2);
    return DS_ERR_ARG;
}

}  // namespace ds

using namespace ds;

extern "C" int64_t ds_gram_scratch_elems(int p, int q) { return gram_scratch_elems(p, q); }

extern "C" int ds_gram_f64(const double* A, int64_t lda, int p, const double* B, int64_t ldb, int q, int64_t n,
                           double* G, int64_t ldg, double* partial, void* stream) {
    return gram_f64(A, lda, p, B, ldb, q, n, G, ldg, partial, (cudaStream_t)stream);
}

extern "C" int ds_block_gemm_f64(const double* A, int64_t lda, int p, const double* C, int64_t ldc, int q, int64_t n,
                                 double beta, double* Y, int64_t ldy, void* stream) {
    return block_gemm_f64(A, lda, p, C, ldc, q, n, 1.0, beta, Y, ldy, (cudaStream_t)stream);
}

extern "C" int ds_rr_update_f64(const double* S, const double* KS, const double* MS, int64_t lda, int prow, int m,
                                const double* C1, const double* C2, int q2, int64_t ldc, int64_t n, double* S_out,
                                double* KS_out, double* MS_out, int64_t ldy, void* stream) {
    const double* A[3] = {S, KS, MS};
    double* Y[3] = {S_out, KS_out, MS_out};
    return rr_update_f64(A, lda, prow, m, C1, C2, q2, ldc, n, Y, ldy, (cudaStream_t)stream);
}
