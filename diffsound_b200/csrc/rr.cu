// Rayleigh-Ritz bookkeeping of the LOBPCG iteration, round 2: only what involves the NEW search block W is
// computed by a pass over the n-sized buffers; everything else follows from the previous step's small matrices.
//
// Reference behaviour replaced: the Gram products S^T A S, S^T B S and the basis update of
// /root/reference/src/lobpcg/_lobpcg.py:433-525 (_update_ortho / _get_rayleigh_ritz_transform), which the reference
// (and round 1 of this library, csrc/gram_sym.cu) recompute in full over [X | W | P] every step.
//
// After a Ritz step with coefficient matrix C (rows = slots of [X | W | P], columns = Ritz rank):
//   X' = S C[:, :m]          X'^T K X' = diag(theta),  X'^T M X' = I        (known)
//   P' = [W P] C[m:, :m]     X'^T A P' = C1^T G C_wp,  P'^T A P' = C_wp^T G C_wp   (m x m products, k_gram_algebra)
// so the next step only needs the strips  (K W)^T [X W P]  and  (M W)^T [X W P]  of its new W (k_gram_strip: 216
// instead of 342 DMMA tile products per 4 rows, 1.58 GB instead of 2.85 GB read), and since X' = X C_x + P' the update
// needs 144 x 48 coefficients per buffer instead of 144 x 48 + 96 x 48 (k_rr_update2: P' first, X' accumulated on top).
// The driver (csrc/lobpcg.cu) recomputes the full Gram pair with k_gram_sym2 every few steps and whenever the small
// Cholesky fails, which bounds the drift of the recurrences.
#include "common.cuh"
#include "../../include/diffsound_sm100.h"
#include "kernels.cuh"
#include "ptx.cuh"
#include <algorithm>

namespace ds {

// ---------------------------------------------------------------------------------------------------------------
// strips: GsK[r][c] = sum_rows KW[row][r] S[row][c],  GsM likewise;  r < wa (16 | 32 | 48), c < ld (<= 144)
// ---------------------------------------------------------------------------------------------------------------
constexpr int ST_THREADS = 512;              // 16 compute warps (+ one producer warp in the PRODUCER variants)
constexpr int ST_WARPS = ST_THREADS / 32;
constexpr int ST_PITCH = 252;              // doubles per staged row: [KW 48 | MW 48 | S 144] = 240, 252 = 12 (mod 16)
constexpr int ST_MAXA = 12, ST_MAXB = 18;  // A-side tiles (K and M strips), B-side tiles
constexpr int ST_WA = 4, ST_WB = 4;        // warp grid
constexpr int ST_TA = 3, ST_TB = 5;        // tiles per warp: 4 x 3 >= 12, 4 x 5 >= 18
constexpr size_t st_smem(int rows, int stages) { return (size_t)stages * rows * ST_PITCH * sizeof(double) + 2 * stages * sizeof(uint64_t); }

// ST_ROWS rows per ring stage, ST_STAGES stages; PRODUCER: a 17th warp issues the bulk copies (the compute warps never wait
// on the `empty` barriers), else warp 0 issues them between its own tile products
template <int ST_ROWS, int ST_STAGES, bool PRODUCER>
__global__ void __launch_bounds__(ST_THREADS + (PRODUCER ? 32 : 0), 1)
k_gram_strip(const double* __restrict__ KW, const double* __restrict__ MW, int64_t ldw, int wa,
             const double* __restrict__ S, int64_t lds, int ncol, int64_t n, double* __restrict__ partial) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* buf = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(buf + (size_t)ST_STAGES * ST_ROWS * ST_PITCH);
    uint64_t* empty = full + ST_STAGES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_chunks = (n + ST_ROWS - 1) / ST_ROWS;
    const int64_t mine = (n_chunks > blockIdx.x) ? (n_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (tid == 0) {
        for (int s = 0; s < ST_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], ST_WARPS); }
        fence_barrier_init();
    }
    __syncthreads();
    const int na_t = 2 * wa / 8, nb_t = ncol / 8;
    const int wi = warp / ST_WB, wj = warp % ST_WB;
    int aoff[ST_TA], boff[ST_TB];            // column offsets inside a staged row, -1 = none
#pragma unroll
    for (int i = 0; i < ST_TA; ++i) { const int t = wi + ST_WA * i; aoff[i] = t < na_t ? 8 * t : -1; }
#pragma unroll
    for (int j = 0; j < ST_TB; ++j) { const int t = wj + ST_WB * j; boff[j] = t < nb_t ? 2 * wa + 8 * t : -1; }
    double acc[ST_TA][ST_TB][2];
#pragma unroll
    for (int i = 0; i < ST_TA; ++i)
#pragma unroll
        for (int j = 0; j < ST_TB; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const uint32_t wbytes = (uint32_t)(wa * sizeof(double)), sbytes = (uint32_t)(ncol * sizeof(double));
    auto issue = [&](int64_t it) {      // warp 0, all lanes
        const int s = (int)(it % ST_STAGES);
        const int64_t r0 = (blockIdx.x + it * (int64_t)gridDim.x) * ST_ROWS;
        const int rows = (int)min((int64_t)ST_ROWS, n - r0);
        if (it >= ST_STAGES) mbar_wait(&empty[s], (uint32_t)(((it / ST_STAGES) - 1) & 1));
        if (rows < ST_ROWS) {
            double* tail = buf + ((size_t)s * ST_ROWS + rows) * ST_PITCH;
            for (int q = lane; q < (ST_ROWS - rows) * ST_PITCH; q += 32) tail[q] = 0.0;
        }
        __syncwarp();
        if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)rows * (2u * wbytes + sbytes));
        __syncwarp();
        if (lane < rows) {
            double* dst = buf + ((size_t)s * ST_ROWS + lane) * ST_PITCH;
            tma_load_1d(dst, KW + (r0 + lane) * ldw, wbytes, &full[s]);
            tma_load_1d(dst + wa, MW + (r0 + lane) * ldw, wbytes, &full[s]);
            tma_load_1d(dst + 2 * wa, S + (r0 + lane) * lds, sbytes, &full[s]);
        }
    };
    if (PRODUCER) {
        if (warp == ST_WARPS) {              // producer warp: the whole ring, nothing else
            for (int64_t it = 0; it < mine; ++it) issue(it);
            return;
        }
    } else if (warp == 0) {
        for (int64_t it = 0; it < min((int64_t)ST_STAGES, mine); ++it) issue(it);
    }

    const int kk = lane & 3, mm = lane >> 2;
    for (int64_t it = 0; it < mine; ++it) {
        const int s = (int)(it % ST_STAGES);
        mbar_wait(&full[s], (uint32_t)((it / ST_STAGES) & 1));
        const double* base = buf + (size_t)s * ST_ROWS * ST_PITCH + kk * ST_PITCH + mm;
#pragma unroll
        for (int k0 = 0; k0 < ST_ROWS; k0 += 4) {
            const double* rp = base + k0 * ST_PITCH;
            double a[ST_TA], b[ST_TB];
#pragma unroll
            for (int i = 0; i < ST_TA; ++i) a[i] = aoff[i] >= 0 ? rp[aoff[i]] : 0.0;
#pragma unroll
            for (int j = 0; j < ST_TB; ++j) b[j] = boff[j] >= 0 ? rp[boff[j]] : 0.0;
#pragma unroll
            for (int i = 0; i < ST_TA; ++i)
#pragma unroll
                for (int j = 0; j < ST_TB; ++j)
                    if (aoff[i] >= 0 && boff[j] >= 0) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (!PRODUCER && warp == 0 && it >= 1 && it - 1 + ST_STAGES < mine) issue(it - 1 + ST_STAGES);
    }
    // partial[cta][at][bt][64]: lane holds C[row = lane>>2][col = 2 (lane&3) + {0,1}]
    double* out = partial + (size_t)blockIdx.x * ST_MAXA * ST_MAXB * 64;
#pragma unroll
    for (int i = 0; i < ST_TA; ++i)
#pragma unroll
        for (int j = 0; j < ST_TB; ++j)
            if (aoff[i] >= 0 && boff[j] >= 0) {
                const int at = wi + ST_WA * i, bt = wj + ST_WB * j;
                reinterpret_cast<double2*>(out + ((size_t)at * ST_MAXB + bt) * 64)[lane] =
                    make_double2(acc[i][j][0], acc[i][j][1]);
            }
}

// Gs[q][r][c] (q = 0: K, 1: M; row-major wa x ldg), one CTA of 64 threads per tile, fixed summation order
__global__ void __launch_bounds__(64)
k_gram_strip_reduce(const double* __restrict__ partial, int nparts, int wa, int nb_t, double* __restrict__ GsK,
                    double* __restrict__ GsM, int64_t ldg) {
    const int at = blockIdx.x / nb_t, bt = blockIdx.x % nb_t;
    const int idx = threadIdx.x;
    double s = 0.0;
    for (int c = 0; c < nparts; ++c) s += partial[((size_t)c * ST_MAXA * ST_MAXB + (size_t)at * ST_MAXB + bt) * 64 + idx];
    const int lane = idx >> 1, half = idx & 1;
    const int r = lane >> 2, col = 2 * (lane & 3) + half;
    const int wt = wa / 8;
    double* G = at < wt ? GsK : GsM;
    const int arow = 8 * (at < wt ? at : at - wt) + r;
    G[(int64_t)arow * ldg + 8 * bt + col] = s;
}

int64_t gram_strip_scratch_elems(int num_sms) { return (int64_t)num_sms * ST_MAXA * ST_MAXB * 64; }

int gram_strip(const double* KW, const double* MW, int64_t ldw, int wa, const double* S, int64_t lds, int ncol, int64_t n,
               double* GsK, double* GsM, int64_t ldg, double* partial, int num_sms, cudaStream_t stream) {
    DS_REQUIRE(KW && MW && S && GsK && GsM && partial, "gram_strip: null argument");
    DS_REQUIRE(wa == 16 || wa == 32 || wa == 48, "gram_strip: wa=%d must be 16, 32 or 48", wa);
    DS_REQUIRE(ncol % 8 == 0 && ncol > 0 && ncol <= 144 && lds >= ncol && ldw >= wa, "gram_strip: bad widths");
    DS_REQUIRE(((uintptr_t)KW % 16 == 0) && ((uintptr_t)MW % 16 == 0) && ((uintptr_t)S % 16 == 0) && ldw % 2 == 0 && lds % 2 == 0,
               "gram_strip: blocks must be 16-byte aligned with even leading dimensions");
    DS_REQUIRE(n > 0, "gram_strip: n must be positive");
    ProfScope prof(PROF_GRAM, stream);
    // Ring geometry measured at n = 823 875, wa = 48 (scripts/bench_dense.py, profiles/r2q_gram_strip_variants.txt):
    // 8 rows x 8 stages, copies issued by warp 0: 1.058 ms (21.5 TFLOP/s); the same with a producer warp: 0.854 ms;
    // 16 rows x 5 stages + producer warp: 0.840 ms (27.1 TFLOP/s = 0.73 of the measured DMMA peak).
    constexpr int ROWS = 16, STAGES = 5;
    constexpr size_t smem = st_smem(ROWS, STAGES);
    static bool attr = false;
    if (!attr) {
        DS_CUDA(cudaFuncSetAttribute(k_gram_strip<ROWS, STAGES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    const int64_t chunks = ceil_div(n, ROWS);
    const int ctas = (int)(chunks < num_sms ? chunks : num_sms);
    k_gram_strip<ROWS, STAGES, true><<<ctas, ST_THREADS + 32, smem, stream>>>(KW, MW, ldw, wa, S, lds, ncol, n, partial);
    DS_LAUNCH_CHECK();
    k_gram_strip_reduce<<<(2 * wa / 8) * (ncol / 8), 64, 0, stream>>>(partial, ctas, wa, ncol / 8, GsK, GsM, ldg);
    DS_LAUNCH_CHECK();
    prof_account(PROF_GRAM, (double)n * (2.0 * wa + ncol) * 8.0, 2.0 * (double)n * (2.0 * wa) * ncol);
    return DS_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// small-matrix algebra on the ldg x ldg Gram pair (one CTA per matrix)
// ---------------------------------------------------------------------------------------------------------------
struct GramAlg {
    const double* G[2];      // current full symmetric Gram matrices (K, M)
    double* Gn[2];           // next
    const double* C;         // coefficient matrix of the Ritz step: rows = slots, columns = rank
    const double* theta;     // Ritz values (ascending)
    int64_t ldg, ldc;
    int m;                   // block size; slots [0, m) X, [m, 2m) W, [2m, 3m) P
};

// T[q] = G[q][:, m:3m] C[m:3m, :m]   (N x m per matrix): grid (N / 16, 2), 16 rows of T per CTA
__global__ void __launch_bounds__(256)
k_gram_alg_t(const __grid_constant__ GramAlg g, double* __restrict__ Tg) {
    extern __shared__ __align__(16) double sm[];
    const int m = g.m, N = 3 * m, q = blockIdx.y, r0 = 16 * blockIdx.x;
    double* Cw = sm;                       // [2m][m]   C[m:3m, :m]
    double* Gr = sm + (size_t)2 * m * m;   // [16][2m]  G[r0 .. r0+16, m:3m]
    const double* __restrict__ G = g.G[q];
    for (int t = threadIdx.x; t < 2 * m * m; t += blockDim.x) Cw[t] = g.C[(int64_t)(m + t / m) * g.ldc + (t % m)];
    for (int t = threadIdx.x; t < 16 * 2 * m; t += blockDim.x)
        Gr[t] = G[(int64_t)(r0 + t / (2 * m)) * g.ldg + m + (t % (2 * m))];
    __syncthreads();
    for (int t = threadIdx.x; t < 16 * m; t += blockDim.x) {
        const int i = t / m, j = t - i * m;
        double s = 0.0;
#pragma unroll 8
        for (int k = 0; k < 2 * m; ++k) s = fma(Gr[i * 2 * m + k], Cw[k * m + j], s);
        Tg[((size_t)q * N + r0 + i) * m + j] = s;
    }
}

// Gn: X'X' = diag(theta) | I;  X'P' = C1^T T;  P'P' = C_wp^T T.  grid (m / 8, 2): 8 columns a of C per CTA.
// Gn must be zero on entry (W rows / columns stay zero until k_gram_insert fills them).
__global__ void __launch_bounds__(256)
k_gram_alg_g(const __grid_constant__ GramAlg g, const double* __restrict__ Tg) {
    extern __shared__ __align__(16) double sm[];
    const int m = g.m, N = 3 * m, q = blockIdx.y, a0 = 8 * blockIdx.x;
    double* T = sm;                        // [N][m]
    double* Ca = sm + (size_t)N * m;       // [N][8]   C[:, a0 .. a0+8]
    double* __restrict__ Gn = g.Gn[q];
    for (int t = threadIdx.x; t < N * m; t += blockDim.x) T[t] = Tg[(size_t)q * N * m + t];
    for (int t = threadIdx.x; t < N * 8; t += blockDim.x) Ca[t] = g.C[(int64_t)(t / 8) * g.ldc + a0 + (t % 8)];
    __syncthreads();
    for (int t = threadIdx.x; t < 8 * m; t += blockDim.x) {
        const int al = t / m, j = t - al * m, a = a0 + al;
        double sx = 0.0, sp = 0.0;
#pragma unroll 8
        for (int i = 0; i < m; ++i) sx = fma(Ca[i * 8 + al], T[i * m + j], sx);
#pragma unroll 8
        for (int i = m; i < N; ++i) sp = fma(Ca[i * 8 + al], T[i * m + j], sp);
        sx += sp;
        Gn[(int64_t)a * g.ldg + 2 * m + j] = sx;
        Gn[(int64_t)(2 * m + j) * g.ldg + a] = sx;
        Gn[(int64_t)(2 * m + a) * g.ldg + 2 * m + j] = sp;       // symmetric up to rounding; consumers read the upper triangle
    }
    if (threadIdx.x < 8) {
        const int a = a0 + threadIdx.x;
        Gn[(int64_t)a * g.ldg + a] = q == 0 ? g.theta[a] : 1.0;
    }
}

// rows / columns [m, 2m) of the Gram pair from the strips of the new W (wa columns in play, the rest zero);
// the W-W block is symmetrised from its upper triangle
__global__ void __launch_bounds__(256)
k_gram_insert(double* __restrict__ GK, double* __restrict__ GM, int64_t ldg, const double* __restrict__ GsK,
              const double* __restrict__ GsM, int64_t lds, int m, int wa, int N) {
    double* G = blockIdx.y == 0 ? GK : GM;
    const double* Gs = blockIdx.y == 0 ? GsK : GsM;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * N) return;
    const int r = t / N, c = t - r * N;          // W column r, any slot c
    double v = 0.0;
    if (r < wa) {
        if (c >= m && c < 2 * m) {
            const int c2 = c - m;
            v = c2 < wa ? (r <= c2 ? Gs[(int64_t)r * lds + c] : Gs[(int64_t)c2 * lds + m + r]) : 0.0;
        } else {
            v = Gs[(int64_t)r * lds + c];
        }
    }
    G[(int64_t)(m + r) * ldg + c] = v;
    G[(int64_t)c * ldg + m + r] = v;
}

// mirror the upper triangle (what k_gram_sym2 writes) into the lower one
__global__ void __launch_bounds__(256)
k_sym_upper(double* __restrict__ GK, double* __restrict__ GM, int64_t ldg, int N) {
    double* G = blockIdx.y == 0 ? GK : GM;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * N) return;
    const int i = t / N, j = t - i * N;
    if (i > j) G[(int64_t)i * ldg + j] = G[(int64_t)j * ldg + i];
}

int64_t gram_algebra_scratch_elems() { return 2 * 144 * 48; }

int gram_algebra(const double* GK, const double* GM, double* GKn, double* GMn, int64_t ldg, const double* C, int64_t ldc,
                 const double* theta, int m, double* scratch, cudaStream_t stream) {
    DS_REQUIRE(GK && GM && GKn && GMn && C && theta && scratch, "gram_algebra: null argument");
    DS_REQUIRE(m == 16 || m == 32 || m == 48, "gram_algebra: m=%d must be 16, 32 or 48", m);
    DS_REQUIRE(ldg >= 3 * m && ldc >= m && GK != GKn && GM != GMn, "gram_algebra: bad leading dimensions / aliasing");
    GramAlg g;
    g.G[0] = GK; g.G[1] = GM; g.Gn[0] = GKn; g.Gn[1] = GMn; g.C = C; g.theta = theta; g.ldg = ldg; g.ldc = ldc; g.m = m;
    const int N = 3 * m;
    ProfScope prof(PROF_EIGH, stream);
    DS_CUDA(cudaMemsetAsync(GKn, 0, sizeof(double) * ldg * N, stream));
    DS_CUDA(cudaMemsetAsync(GMn, 0, sizeof(double) * ldg * N, stream));
    const size_t sm1 = ((size_t)2 * m * m + 16 * 2 * m) * sizeof(double), sm2 = ((size_t)N * m + N * 8) * sizeof(double);
    static bool attr = false;
    if (!attr) {
        DS_CUDA(cudaFuncSetAttribute(k_gram_alg_t, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((2 * 48 * 48 + 16 * 96) * sizeof(double))));
        DS_CUDA(cudaFuncSetAttribute(k_gram_alg_g, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((144 * 48 + 144 * 8) * sizeof(double))));
        attr = true;
    }
    k_gram_alg_t<<<dim3(N / 16, 2), 256, sm1, stream>>>(g, scratch);
    DS_LAUNCH_CHECK();
    k_gram_alg_g<<<dim3(m / 8, 2), 256, sm2, stream>>>(g, scratch);
    DS_LAUNCH_CHECK();
    return DS_OK;
}

int gram_insert(double* GK, double* GM, int64_t ldg, const double* GsK, const double* GsM, int64_t lds, int m, int wa,
                cudaStream_t stream) {
    DS_REQUIRE(GK && GM && GsK && GsM && wa <= m && ldg >= 3 * m && lds >= 3 * m, "gram_insert: bad argument");
    ProfScope prof(PROF_EIGH, stream);
    k_gram_insert<<<dim3((unsigned)ceil_div((int64_t)m * 3 * m, 256), 2), 256, 0, stream>>>(GK, GM, ldg, GsK, GsM, lds, m, wa, 3 * m);
    DS_LAUNCH_CHECK();
    return DS_OK;
}

int sym_upper(double* GK, double* GM, int64_t ldg, int N, cudaStream_t stream) {
    ProfScope prof(PROF_EIGH, stream);
    k_sym_upper<<<dim3((unsigned)ceil_div((int64_t)N * N, 256), 2), 256, 0, stream>>>(GK, GM, ldg, N);
    DS_LAUNCH_CHECK();
    return DS_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// update:  for A in {S, KS, MS}:  Y = A[:, wlo:whi) C[wlo:whi, :m] (+ A[:, 2m:3m) C[2m:3m, :m] when use_p);
//          Anew[:, 2m:3m) = Y;  Anew[:, 0:m) = Y + A[:, 0:m) C[0:m, :m]
// ---------------------------------------------------------------------------------------------------------------
constexpr int U2_THREADS = 256;

__host__ __device__ inline int pad4mod16(int q) {          // smallest s >= q with s % 16 == 4 (conflict-free B fragments)
    int s = (q / 16) * 16 + 4;
    return s >= q ? s : s + 16;
}

struct RRUpdate2Args {
    const double* A[3];
    double* Y[3];
    int64_t lda, ldy, n;
    const double* C;
    int64_t ldc;
    int m, wa, use_p;
};

template <int QT>
__global__ void __launch_bounds__(U2_THREADS)
k_rr_update2(const __grid_constant__ RRUpdate2Args g) {
    extern __shared__ __align__(16) double Cs[];     // [3m][qs]
    constexpr int q = 8 * QT, NS = 2;
    const int qs = pad4mod16(q), m = g.m, N = 3 * m;
    for (int t = threadIdx.x; t < N * q; t += blockDim.x) {
        const int r = t / q, c = t - r * q;
        Cs[r * qs + c] = g.C[(int64_t)r * g.ldc + c];
    }
    __syncthreads();
    const double* __restrict__ A = g.A[blockIdx.y];
    double* __restrict__ Y = g.Y[blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kk = lane & 3, mm = lane >> 2;
    const int64_t n_strips = (g.n + 7) / 8;
    // 16-column groups in processing order: W, P (then P' is stored), X (then X' is stored)
    const int ngw = g.wa / 16, ngp = g.use_p ? m / 16 : 0, ngx = m / 16, ng = ngw + ngp + ngx;
    auto k0_of = [&](int gi) { return gi < ngw ? m + 16 * gi : (gi < ngw + ngp ? 2 * m + 16 * (gi - ngw) : 16 * (gi - ngw - ngp)); };
    for (int64_t strip = ((int64_t)blockIdx.x * (U2_THREADS / 32) + warp) * NS; strip < n_strips;
         strip += (int64_t)gridDim.x * (U2_THREADS / 32) * NS) {
        bool ok[NS];
        const double* ap[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int64_t row = (strip + s) * 8 + mm;
            ok[s] = row < g.n;
            ap[s] = A + (ok[s] ? row : 0) * g.lda + kk;
        }
        double acc[NS][QT][2];
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int t = 0; t < QT; ++t) acc[s][t][0] = acc[s][t][1] = 0.0;
        double a[NS][4], an[NS][4];
        {
            const int k0 = k0_of(0);
#pragma unroll
            for (int s = 0; s < NS; ++s)
#pragma unroll
                for (int u = 0; u < 4; ++u) a[s][u] = ok[s] ? __ldg(ap[s] + k0 + 4 * u) : 0.0;
        }
#pragma unroll 1
        for (int gi = 0; gi < ng; ++gi) {
            const int k0 = k0_of(gi);
            if (gi + 1 < ng) {
                const int k1 = k0_of(gi + 1);
#pragma unroll
                for (int s = 0; s < NS; ++s)
#pragma unroll
                    for (int u = 0; u < 4; ++u) an[s][u] = ok[s] ? __ldg(ap[s] + k1 + 4 * u) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double* bp = Cs + (k0 + 4 * u + kk) * qs + mm;
#pragma unroll
                for (int t = 0; t < QT; ++t) {
                    const double b = bp[8 * t];
#pragma unroll
                    for (int s = 0; s < NS; ++s) dmma_m8n8k4(acc[s][t][0], acc[s][t][1], a[s][u], b);
                }
            }
            if (gi == ngw + ngp - 1 || gi == ng - 1) {
                const int off = gi == ng - 1 ? 0 : 2 * m;      // P' after the W and P groups, X' after the X groups
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    if (!ok[s]) continue;
                    double* yp = Y + ((strip + s) * 8 + mm) * g.ldy + off + 2 * kk;
#pragma unroll
                    for (int t = 0; t < QT; ++t)
                        *reinterpret_cast<double2*>(yp + 8 * t) = make_double2(acc[s][t][0], acc[s][t][1]);
                }
            }
#pragma unroll
            for (int s = 0; s < NS; ++s)
#pragma unroll
                for (int u = 0; u < 4; ++u) a[s][u] = an[s][u];
        }
    }
}

int rr_update2_f64(const double* const A[3], int64_t lda, int m, int wa, int use_p, const double* C, int64_t ldc, int64_t n,
                   double* const Y[3], int64_t ldy, cudaStream_t stream) {
    DS_REQUIRE(m == 16 || m == 32 || m == 48, "rr_update2: m=%d must be 16, 32 or 48", m);
    DS_REQUIRE(wa >= 16 && wa % 16 == 0 && wa <= m && C && ldc >= m && lda >= 3 * m && ldy >= 3 * m, "rr_update2: bad argument");
    RRUpdate2Args g;
    for (int b = 0; b < 3; ++b) {
        DS_REQUIRE(A[b] && Y[b] && A[b] != Y[b] && (uintptr_t)Y[b] % 16 == 0, "rr_update2: bad buffer %d", b);
        g.A[b] = A[b];
        g.Y[b] = Y[b];
    }
    DS_REQUIRE(ldy % 2 == 0, "rr_update2: ldy must be even");
    g.lda = lda; g.ldy = ldy; g.n = n; g.C = C; g.ldc = ldc; g.m = m; g.wa = wa; g.use_p = use_p;
    ProfScope prof(PROF_GEMM, stream);
    const size_t smem = (size_t)3 * m * pad4mod16(m) * sizeof(double);
    const int64_t strips = (n + 7) / 8;
    const int per_buf = (int)std::min<int64_t>((strips + 15) / 16, 148 * 2);
    auto launch = [&](auto kern) -> int {
        DS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<dim3(per_buf, 3), U2_THREADS, smem, stream>>>(g);
        DS_LAUNCH_CHECK();
        return DS_OK;
    };
    const double kin = m + wa + (use_p ? m : 0);
    prof_account(PROF_GEMM, 3.0 * (double)n * (kin + 2.0 * m) * 8.0, 3.0 * 2.0 * (double)n * kin * m);
    if (m == 16) return launch(k_rr_update2<2>);
    if (m == 32) return launch(k_rr_update2<4>);
    return launch(k_rr_update2<6>);
}

// ---------------------------------------------------------------------------------------------------------------
// FP64 peak microbenchmarks (register-resident): mode 0 = DFMA, mode 1 = DMMA m8n8k4
// ---------------------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(512)
k_fp64_peak(int iters, double* __restrict__ out) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 1e-3 * (threadIdx.x + i);
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * (blockIdx.x + 1);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma_m8n8k4(acc[2 * i], acc[2 * i + 1], a, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;      // keep the loop alive
}

}  // namespace ds

using namespace ds;

extern "C" int ds_fp64_peak(int mode, int iters, int ctas_per_sm, double* scratch, double* tflops_host, void* stream) {
    DS_REQUIRE((mode == 0 || mode == 1) && iters > 0 && ctas_per_sm > 0 && scratch && tflops_host, "ds_fp64_peak: bad argument");
    int dev = 0, sms = 0;
    DS_CUDA(cudaGetDevice(&dev));
    DS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    DS_CUDA(cudaEventCreate(&e0));
    DS_CUDA(cudaEventCreate(&e1));
    const int grid = sms * ctas_per_sm;
    for (int rep = 0; rep < 2; ++rep) {          // first pass: warm-up
        DS_CUDA(cudaEventRecord(e0, st));
        if (mode == 0) k_fp64_peak<0><<<grid, 512, 0, st>>>(iters, scratch);
        else k_fp64_peak<1><<<grid, 512, 0, st>>>(iters, scratch);
        DS_LAUNCH_CHECK();
        DS_CUDA(cudaEventRecord(e1, st));
        DS_CUDA(cudaEventSynchronize(e1));
    }
    float ms = 0.f;
    DS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    // DFMA: 16 fma x 2 flop per thread-iteration; DMMA: 8 mma x (8 x 8 x 4 x 2 = 512 flop) per warp-iteration
    const double flops = mode == 0 ? (double)grid * 512 * iters * 32.0 : (double)grid * 16 * iters * 8.0 * 512.0;
    *tflops_host = flops / (ms * 1e-3) / 1e12;
    return DS_OK;
}

extern "C" int ds_gram_strip_f64(const double* KW, const double* MW, int64_t ldw, int wa, const double* S, int64_t lds,
                                 int ncol, int64_t n, double* GsK, double* GsM, int64_t ldg, double* partial, void* stream) {
    int dev = 0, sms = 0;
    DS_CUDA(cudaGetDevice(&dev));
    DS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    return gram_strip(KW, MW, ldw, wa, S, lds, ncol, n, GsK, GsM, ldg, partial, sms, (cudaStream_t)stream);
}

extern "C" int64_t ds_gram_strip_scratch_elems(void) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return -1;
    return gram_strip_scratch_elems(sms);
}

extern "C" int ds_rr_update2_f64(const double* S, const double* KS, const double* MS, int64_t lda, int m, int wa, int use_p,
                                 const double* C, int64_t ldc, int64_t n, double* S_out, double* KS_out, double* MS_out,
                                 int64_t ldy, void* stream) {
    const double* A[3] = {S, KS, MS};
    double* Y[3] = {S_out, KS_out, MS_out};
    return rr_update2_f64(A, lda, m, wa, use_p, C, ldc, n, Y, ldy, (cudaStream_t)stream);
}

extern "C" int64_t ds_gram_algebra_scratch_elems(void) { return gram_algebra_scratch_elems(); }

extern "C" int ds_gram_algebra_f64(const double* GK, const double* GM, double* GKn, double* GMn, int64_t ldg, const double* C,
                                   int64_t ldc, const double* theta, int m, double* scratch, void* stream) {
    return gram_algebra(GK, GM, GKn, GMn, ldg, C, ldc, theta, m, scratch, (cudaStream_t)stream);
}
