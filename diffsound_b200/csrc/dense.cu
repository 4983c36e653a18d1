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

__global__ void k_gram_reduce(const double* __restrict__ partial, int nparts, int p, int q, double* __restrict__ G,
                              int64_t ldg) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p * q) return;
    double s = 0.0;
    for (int c = 0; c < nparts; ++c) s += partial[(size_t)c * p * q + idx];
    G[(int64_t)(idx / q) * ldg + (idx % q)] = s;
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
    k_gram_reduce<<<(p * q + 255) / 256, 256, 0, stream>>>(partial, ctas, p, q, G, ldg);
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
