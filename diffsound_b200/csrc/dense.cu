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
//  * eigh_generalized_f64:  one CTA; Cholesky of GM and GK + sigma GM, then
//                one-sided Jacobi on the rows of L^-1 R in shared memory.
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

// ---------------------------------------------------------------------------
// Generalised symmetric eigenproblem, one CTA
// ---------------------------------------------------------------------------
constexpr int EIG_MAXN = 144;
constexpr int JG = 8;                       // lanes per Jacobi pair group
constexpr int JE = EIG_MAXN / JG;           // row elements per lane
constexpr int EIG_THREADS = (EIG_MAXN / 2) * JG;   // 288: one group per pair of line positions
constexpr int JXS = 152;                    // mailbox slot pitch in doubles: 8 (mod 16) -> the 4 groups of a warp tile the banks
constexpr int JXN = EIG_MAXN;               // slot[JXN] = |row|^2, slot[JXN + 1] = scale

// in-place Cholesky (lower) of the N x N matrix in shared memory (row stride ld).
// returns 0 or failing column + 1 (same value in every thread).
__device__ int chol_lower(double* S, int N, int ld, int* s_flag) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int k = 0; k < N; ++k) {
        if (tid == 0) {
            double d = S[k * ld + k];
            if (!(d > 0.0)) *s_flag = k + 1;
            else S[k * ld + k] = sqrt(d);
        }
        __syncthreads();
        if (*s_flag) return *s_flag;
        double dk = S[k * ld + k];
        for (int i = k + 1 + tid; i < N; i += nt) S[i * ld + k] /= dk;
        __syncthreads();
        int m = N - k - 1;
        for (int t = tid; t < m * m; t += nt) {
            int i = k + 1 + t / m, j = k + 1 + t % m;
            if (j <= i) S[i * ld + j] -= S[i * ld + k] * S[j * ld + k];
        }
        __syncthreads();
    }
    return 0;
}

struct EigIdx {
    short v[EIG_MAXN];
};

// upper-triangle read through the slot map: entry (i, j) of the compact problem
__device__ __forceinline__ double g_up(const double* __restrict__ G, int64_t ldg, const EigIdx& ix, int i, int j) {
    int a = ix.v[i], b = ix.v[j];
    return a <= b ? G[(int64_t)a * ldg + b] : G[(int64_t)b * ldg + a];
}

__global__ void __launch_bounds__(EIG_THREADS)
k_eigh_generalized(const double* __restrict__ GK, const double* __restrict__ GM, int N, int64_t ldg,
                   const __grid_constant__ EigIdx ix, double sigma_in, double* __restrict__ theta,
                   double* __restrict__ C, int64_t ldc, double* __restrict__ scratch, int* __restrict__ info) {
    extern __shared__ __align__(16) double S[];  // [N][ld]
    const int ld = N + 2;
    // the Jacobi mailbox ([(N+1)/2 + 1][JXS]) later aliases the matrix; the small arrays sit behind both
    const size_t mat_elems = (size_t)N * ld, mbox_elems = (size_t)((N + 1) / 2 + 1) * JXS;
    double* s_scale = S + (mat_elems > mbox_elems ? mat_elems : mbox_elems);   // [N]
    double* s_theta = s_scale + N;           // [N]
    int* s_rank = reinterpret_cast<int*>(s_theta + N);  // [N]
    int* s_flag = s_rank + N;                // [2]
    const int tid = threadIdx.x, nt = blockDim.x;
    double* Lg = scratch;                    // L, row-major [N][N]
    double* Rt = scratch + (size_t)N * N;    // R^T, row-major: Rt[k][i] = R[i][k]
    if (tid == 0) { s_flag[0] = 0; s_flag[1] = 0; }
    for (int i = tid; i < N; i += nt) {
        double d = g_up(GM, ldg, ix, i, i);
        s_scale[i] = d > 0.0 ? rsqrt(d) : 1.0;
    }
    __syncthreads();
    // sigma < 0: automatic shift = |sigma| * mean diagonal of the scaled GK
    double sigma = sigma_in;
    if (sigma_in < 0.0) {
        double tr = 0.0;
        for (int i = 0; i < N; ++i) tr += fabs(g_up(GK, ldg, ix, i, i)) * s_scale[i] * s_scale[i];
        sigma = -sigma_in * tr / N;
    }
    // ---- L = chol(D GM D)
    for (int t = tid; t < N * N; t += nt) {
        int i = t / N, j = t % N;
        double v = (j <= i) ? g_up(GM, ldg, ix, i, j) : 0.0;
        S[i * ld + j] = v * s_scale[i] * s_scale[j];
    }
    __syncthreads();
    int bad = chol_lower(S, N, ld, s_flag);
    if (bad) { if (tid == 0) { info[0] = bad; info[1] = 0; } return; }
    for (int t = tid; t < N * N; t += nt) {
        int i = t / N, j = t % N;
        Lg[t] = (j <= i) ? S[i * ld + j] : 0.0;
    }
    __syncthreads();
    // ---- R = chol(D (GK + sigma GM) D)
    for (int t = tid; t < N * N; t += nt) {
        int i = t / N, j = t % N;
        double v = 0.0;
        if (j <= i) v = g_up(GK, ldg, ix, i, j) + sigma * g_up(GM, ldg, ix, i, j);
        S[i * ld + j] = v * s_scale[i] * s_scale[j];
    }
    __syncthreads();
    bad = chol_lower(S, N, ld, s_flag);
    if (bad) { if (tid == 0) { info[0] = 1000 + bad; info[1] = 0; } return; }
    for (int t = tid; t < N * N; t += nt) {
        int i = t / N, k = t % N;   // Rt[k][i] = R[i][k]
        Rt[(size_t)k * N + i] = (k <= i) ? S[i * ld + k] : 0.0;
    }
    for (int t = tid; t < N * N; t += nt) {
        int i = t / N, j = t % N;
        if (j > i) S[i * ld + j] = 0.0;
    }
    __threadfence_block();
    __syncthreads();
    // ---- Y = L^-1 R  (forward substitution, one thread per column j; Y lower triangular)
    if (tid < N) {
        int j = tid;
        for (int i = j; i < N; ++i) {
            double v = S[i * ld + j];
            const double* Li = Lg + (size_t)i * N;
            for (int k = j; k < i; ++k) v -= Li[k] * S[k * ld + j];
            S[i * ld + j] = v / Li[i];
        }
    }
    __syncthreads();
    // ---- one-sided Jacobi on the rows y_0..y_{N-1} of Y: converges to Y_final with orthogonal rows,
    //      |y_j|^2 = theta_j + sigma.
    // The rows live in REGISTERS for the whole iteration: group g (4 lanes, 36 elements per lane and
    // row) owns line positions 2g and 2g+1.  Pairs follow the odd-even transposition ordering: even
    // steps rotate positions (2g, 2g+1), odd steps (2g+1, 2g+2), the two rows trade places after every
    // step, so after Np steps every pair has met once and only ONE row per group crosses to a
    // neighbour per step (through a shared-memory mailbox that aliases the no longer needed matrix).
    // Rotations are self-scaling (a' = a - t1 b, b' = b + t2 a with the cosine folded into a per-row
    // scale): 2 DFMA per element instead of 4; |row|^2 is updated by +-t*gamma and recomputed every sweep.
    const int Np = (N + 1) & ~1;          // even number of line positions (an odd N gets one zero row)
    const int G = Np / 2;
    const int grp = tid / JG, gl = tid % JG;
    const bool active = grp < G;
    const double tol = 1.2e-16 * sqrt((double)N);
    double ra[JE], rb[JE];                // even configuration: ra = position 2g, rb = position 2g+1
    double sa = 1.0, sb = 1.0, na = 0.0, nb = 0.0;
#pragma unroll
    for (int i = 0; i < JE; ++i) {
        const int e = gl + JG * i;
        const int p0 = 2 * grp, p1 = 2 * grp + 1;
        ra[i] = (active && p0 < N && e < N) ? S[(size_t)p0 * ld + e] : 0.0;
        rb[i] = (active && p1 < N && e < N) ? S[(size_t)p1 * ld + e] : 0.0;
    }
    __syncthreads();
    double* xfer = S;                     // [G + 1][JXS]: elements, [JXN] = |row|^2, [JXN + 1] = scale
    for (int t = tid; t < JXS; t += nt) xfer[(size_t)G * JXS + t] = 0.0;      // beyond-the-end partner: a zero row
    __syncthreads();
    auto group_sum = [&](double v) {
#pragma unroll
        for (int o = 1; o < JG; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    };
    // rotate (x: scale sx, norm nx) against (y: sy, ny); returns 1 if a rotation was applied
    // (every thread of the CTA calls this: the group sum is a full-warp shuffle)
    auto rotate = [&](bool enable, double (&x)[JE], double& sx, double& nx, double (&y)[JE], double& sy,
                      double& ny) -> int {
        double g0 = 0.0;
#pragma unroll
        for (int i = 0; i < JE; ++i) g0 = fma(x[i], y[i], g0);
        const double ga = group_sum(g0) * sx * sy;
        if (!(enable && nx > 0.0 && ny > 0.0 && fabs(ga) > tol * sqrt(nx * ny))) return 0;
        const double zeta = (ny - nx) / (2.0 * ga);
        const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = rsqrt(1.0 + t * t);
        const double t1 = t * sy / sx, t2 = t * sx / sy;
#pragma unroll
        for (int i = 0; i < JE; ++i) {
            const double xv = x[i], yv = y[i];
            x[i] = fma(-t1, yv, xv);
            y[i] = fma(t2, xv, yv);
        }
        sx *= c; sy *= c;
        nx -= t * ga; ny += t * ga;
        return 1;
    };
    auto send = [&](const double (&x)[JE], double sx, double nx, int slot) {
        double* d = xfer + (size_t)slot * JXS;
#pragma unroll
        for (int i = 0; i < JE; ++i) d[gl + JG * i] = x[i];
        if (gl == 0) { d[JXN] = nx; d[JXN + 1] = sx; }
    };
    auto recv = [&](double (&x)[JE], double& sx, double& nx, int slot) {
        const double* d = xfer + (size_t)slot * JXS;
#pragma unroll
        for (int i = 0; i < JE; ++i) x[i] = d[gl + JG * i];
        nx = d[JXN]; sx = d[JXN + 1];
    };
    int sweep = 0;
    for (; sweep < 24; ++sweep) {
        if (tid == 0) s_flag[1] = 0;
        // fold the scales into the rows and refresh the norms
        {
            double qa = 0.0, qb = 0.0;
#pragma unroll
            for (int i = 0; i < JE; ++i) {
                ra[i] *= sa; rb[i] *= sb;
                qa = fma(ra[i], ra[i], qa); qb = fma(rb[i], rb[i], qb);
            }
            sa = sb = 1.0;
            na = group_sum(qa); nb = group_sum(qb);
        }
        __syncthreads();
        int rotated = 0;
        for (int step = 0; step < Np; step += 2) {
            // even step: positions (2g, 2g+1) = (ra, rb); afterwards the rows trade places, i.e.
            // position 2g is in rb and position 2g+1 in ra
            rotated |= rotate(active, ra, sa, na, rb, sb, nb);
            // odd step: position 2g (in rb) goes to the left neighbour, position 2g+2 arrives in rb
            if (active) send(rb, sb, nb, grp);
            __syncthreads();
            if (active) recv(rb, sb, nb, grp + 1);
            rotated |= rotate(active && grp < G - 1, ra, sa, na, rb, sb, nb);           // (2g+1, 2g+2)
            // trade places: position 2g+1 is now in rb, position 2g+2 in ra -- except for the last group,
            // whose partner is the beyond-the-end zero row: it keeps its row at position 2g+1
            if (active && grp == G - 1) {
#pragma unroll
                for (int i = 0; i < JE; ++i) { const double tv = ra[i]; ra[i] = rb[i]; rb[i] = tv; }
                double tv = sa; sa = sb; sb = tv;
                tv = na; na = nb; nb = tv;
            }
            // next even step: position 2g+2 (in ra) goes to the right neighbour, position 2g arrives in ra
            if (active) send(ra, sa, na, grp + 1);
            __syncthreads();
            if (active) recv(ra, sa, na, grp);
        }
        if (rotated && gl == 0) s_flag[1] = 1;
        __syncthreads();
        const int again = s_flag[1];
        __syncthreads();
        if (!again) break;
    }
    // ---- rows back to shared memory (true values), zero rows of the padding dropped
    if (tid == 0) s_flag[1] = Np;
    __syncthreads();
    if (active && gl == 0) {
        if (na == 0.0) s_flag[1] = 2 * grp;              // at most one zero row exists (odd N)
        if (nb == 0.0) s_flag[1] = 2 * grp + 1;
    }
    __syncthreads();
    {
        const int ph = s_flag[1];
        // the mailbox aliases S: park the rows in registers until everybody has read its last message
        __syncthreads();
        if (active) {
            const int p0 = 2 * grp, p1 = 2 * grp + 1;
            const int d0 = p0 - (p0 > ph ? 1 : 0), d1 = p1 - (p1 > ph ? 1 : 0);
#pragma unroll
            for (int i = 0; i < JE; ++i) {
                const int e = gl + JG * i;
                if (e < N) {
                    if (p0 != ph && d0 < N) S[(size_t)d0 * ld + e] = ra[i] * sa;
                    if (p1 != ph && d1 < N) S[(size_t)d1 * ld + e] = rb[i] * sb;
                }
            }
        }
    }
    __syncthreads();
    // ---- eigenvalues and ascending rank
    for (int j = tid / 32; j < N; j += nt / 32) {
        double v = 0.0;
        for (int e = tid & 31; e < N; e += 32) { double a = S[(size_t)j * ld + e]; v = fma(a, a, v); }
        v = warp_sum(v);
        if ((tid & 31) == 0) s_theta[j] = v - sigma;
    }
    __syncthreads();
    for (int j = tid; j < N; j += nt) {
        double tj = s_theta[j];
        int rk = 0;
        for (int i = 0; i < N; ++i) {
            double ti = s_theta[i];
            rk += (ti < tj) || (ti == tj && i < j);
        }
        s_rank[j] = rk;
        theta[rk] = tj;
    }
    __syncthreads();
    // ---- c_j^T = y_j R^-1 (back substitution, thread per vector), scaled, placed in column rank_j
    if (tid < N) {
        int j = tid;
        double* y = S + (size_t)j * ld;
        for (int k = N - 1; k >= 0; --k) {
            double v = y[k];
            const double* Rk = Rt + (size_t)k * N;   // Rk[i] = R[i][k]
            for (int i = k + 1; i < N; ++i) v -= y[i] * Rk[i];
            y[k] = v / Rk[k];
        }
    }
    __syncthreads();
    for (int t = tid; t < N * N; t += nt) {
        int k = t / N, j = t % N;
        C[(int64_t)ix.v[k] * ldc + s_rank[j]] = S[(size_t)j * ld + k] * s_scale[k];
    }
    if (tid == 0) { info[0] = 0; info[1] = sweep + 1; }
}

int eigh_generalized_f64(const double* GK, const double* GM, int N, int64_t ldg, const int* idx_host, double sigma,
                         double* theta, double* C, int64_t ldc, double* scratch, int* info, cudaStream_t stream) {
    DS_REQUIRE(N >= 2 && N <= EIG_MAXN, "eigh_generalized: N=%d must be in [2,%d]", N, EIG_MAXN);
    DS_REQUIRE(GK && GM && theta && C && scratch && info, "eigh_generalized: null argument");
    ProfScope prof(PROF_EIGH, stream);
    EigIdx ix;
    for (int i = 0; i < EIG_MAXN; ++i) ix.v[i] = (short)(i < N ? (idx_host ? idx_host[i] : i) : 0);
    const size_t mat = (size_t)N * (N + 2), mailbox = (size_t)(((N + 1) / 2) + 1) * JXS;
    size_t smem = ((mat > mailbox ? mat : mailbox) + 2 * N) * sizeof(double) + (N + 4) * sizeof(int);
    DS_CUDA(cudaFuncSetAttribute(k_eigh_generalized, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_eigh_generalized<<<1, EIG_THREADS, smem, stream>>>(GK, GM, N, ldg, ix, sigma, theta, C, ldc, scratch, info);
    DS_LAUNCH_CHECK();
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

extern "C" int ds_eigh_generalized_f64(const double* GK, const double* GM, int N, int64_t ldg, double sigma,
                                       double* theta, double* C, int64_t ldc, double* scratch, int* info,
                                       void* stream) {
    return eigh_generalized_f64(GK, GM, N, ldg, nullptr, sigma, theta, C, ldc, scratch, info, (cudaStream_t)stream);
}
