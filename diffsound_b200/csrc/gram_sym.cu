// Fused symmetric Gram pair of the Rayleigh-Ritz step:  GK = S^T (K S),  GM = S^T (M S)  in ONE pass
// over the three n x ld blocks S, KS, MS (ld <= 144 columns = [X | W | P]).
//
// Reference behaviour replaced: the qform / matmul products S^T A S and S^T B S of
// /root/reference/src/lobpcg/_lobpcg.py:479-525 (_linalg_utils.py:63-73), which the first version of
// this library evaluated as twelve separate 48 x 48 block Grams per iteration (twelve passes).
//
// Design: one persistent CTA per SM (512 threads).  Rows are streamed in chunks of GS_ROWS through a
// GS_STAGES-deep shared-memory ring by per-row 1-D TMA bulk copies (mbarrier full/empty).  Only the
// upper triangle of 8 x 8 tiles over the ACTIVE tile columns is computed (<= 171 tile pairs); every
// warp owns up to GS_EPW tile pairs and keeps both accumulators (K and M) of each in registers for
// the whole kernel: 2 x DMMA m8n8k4 per pair and k-step, the A fragment shared by the two.
// Row pitch = 3 * 148 doubles = 12 (mod 16) words, so the four k-rows of a fragment load fall into
// disjoint bank groups (conflict-free 64-bit loads).  Per-CTA partial tiles are reduced in a fixed
// order by a second kernel (deterministic, no atomics).
#include "common.cuh"
#include "../../include/diffsound_sm100.h"
#include "kernels.cuh"
#include "ptx.cuh"

namespace ds {

constexpr int GS_ROWS = 8;
constexpr int GS_STAGES = 6;
constexpr int GS_THREADS = 512;
constexpr int GS_WARPS = GS_THREADS / 32;
constexpr int GS_SEG = 148;               // doubles per segment (144 used); 148 * 8 B keeps 16-byte alignment
constexpr int GS_PITCH = 3 * GS_SEG;      // 444 = 12 (mod 16)
constexpr int GS_EPW = 11;                // tile pairs per warp: 16 * 11 = 176 >= 171
constexpr int GS_MAX_ENTRIES = GS_WARPS * GS_EPW;
constexpr size_t GS_SMEM = (size_t)GS_STAGES * GS_ROWS * GS_PITCH * sizeof(double) + 2 * GS_STAGES * sizeof(uint64_t);

struct GramPlan {
    unsigned char ti[GS_MAX_ENTRIES];     // tile column of the A side (S), in units of 8 columns
    unsigned char tj[GS_MAX_ENTRIES];     // tile column of the B side (KS / MS), ti <= tj
    int n_entries;
};

__global__ void __launch_bounds__(GS_THREADS, 1)
k_gram_sym2(const double* __restrict__ S, const double* __restrict__ KS, const double* __restrict__ MS, int64_t ld,
            int width, int64_t n, const __grid_constant__ GramPlan plan, double* __restrict__ partial) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* buf = reinterpret_cast<double*>(smem_raw);                           // [STAGES][ROWS][PITCH]
    uint64_t* full = reinterpret_cast<uint64_t*>(buf + (size_t)GS_STAGES * GS_ROWS * GS_PITCH);
    uint64_t* empty = full + GS_STAGES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_chunks = (n + GS_ROWS - 1) / GS_ROWS;
    const int64_t mine = (n_chunks > blockIdx.x) ? (n_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (tid == 0) {
        for (int s = 0; s < GS_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], GS_WARPS);
        }
        fence_barrier_init();
    }
    __syncthreads();

    // entries of this warp: [warp * EPW, warp * EPW + EPW) clipped to n_entries
    int pk[GS_EPW];                     // (8 * ti) << 16 | (8 * tj), -1 = no entry
    const int e0 = warp * GS_EPW;
#pragma unroll
    for (int t = 0; t < GS_EPW; ++t)
        pk[t] = (e0 + t < plan.n_entries) ? ((8 * (int)plan.ti[e0 + t]) << 16 | (8 * (int)plan.tj[e0 + t])) : -1;
    double acc[GS_EPW][2][2];
#pragma unroll
    for (int t = 0; t < GS_EPW; ++t)
#pragma unroll
        for (int q = 0; q < 2; ++q) acc[t][q][0] = acc[t][q][1] = 0.0;

    const uint32_t row_bytes = (uint32_t)(width * sizeof(double));
    auto issue = [&](int64_t it) {      // warp 0, all lanes
        const int s = (int)(it % GS_STAGES);
        const int64_t r0 = (blockIdx.x + it * (int64_t)gridDim.x) * GS_ROWS;
        const int rows = (int)min((int64_t)GS_ROWS, n - r0);
        if (it >= GS_STAGES) mbar_wait(&empty[s], (uint32_t)(((it / GS_STAGES) - 1) & 1));
        if (rows < GS_ROWS) {           // the one short chunk of the matrix: zero the rows TMA will not write
            double* tail = buf + ((size_t)s * GS_ROWS + rows) * GS_PITCH;
            for (int q = lane; q < (GS_ROWS - rows) * GS_PITCH; q += 32) tail[q] = 0.0;
        }
        __syncwarp();
        if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)rows * 3u * row_bytes);
        __syncwarp();
        if (lane < rows) {
            double* dst = buf + ((size_t)s * GS_ROWS + lane) * GS_PITCH;
            tma_load_1d(dst, S + (r0 + lane) * ld, row_bytes, &full[s]);
            tma_load_1d(dst + GS_SEG, KS + (r0 + lane) * ld, row_bytes, &full[s]);
            tma_load_1d(dst + 2 * GS_SEG, MS + (r0 + lane) * ld, row_bytes, &full[s]);
        }
    };
    if (warp == 0)
        for (int64_t it = 0; it < min((int64_t)GS_STAGES, mine); ++it) issue(it);

    const int kk = lane & 3, mm = lane >> 2;
    for (int64_t it = 0; it < mine; ++it) {
        const int s = (int)(it % GS_STAGES);
        mbar_wait(&full[s], (uint32_t)((it / GS_STAGES) & 1));
        const double* base = buf + (size_t)s * GS_ROWS * GS_PITCH + kk * GS_PITCH + mm;
#pragma unroll 1
        for (int k0 = 0; k0 < GS_ROWS; k0 += 4) {
            const double* rp = base + k0 * GS_PITCH;
            int iprev = -2;
            double a = 0.0;
#pragma unroll
            for (int t = 0; t < GS_EPW; ++t) {
                if (pk[t] < 0) continue;
                const int oi = pk[t] >> 16, oj = pk[t] & 0xffff;
                if (oi != iprev) {
                    a = rp[oi];
                    iprev = oi;
                }
                const double bk = rp[GS_SEG + oj];
                const double bm = rp[2 * GS_SEG + oj];
                dmma_m8n8k4(acc[t][0][0], acc[t][0][1], a, bk);
                dmma_m8n8k4(acc[t][1][0], acc[t][1][1], a, bm);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        // refill the stage of the PREVIOUS chunk: by now every warp has normally released it, so the
        // issuing warp (a consumer itself) does not stall the pipeline waiting for the slowest one
        if (warp == 0 && it >= 1 && it - 1 + GS_STAGES < mine) issue(it - 1 + GS_STAGES);
    }
    // partial[cta][entry][matrix][64]; lane holds C[row = lane>>2][col = 2*(lane&3) + {0,1}]
    double* out = partial + (size_t)blockIdx.x * GS_MAX_ENTRIES * 128;
#pragma unroll
    for (int t = 0; t < GS_EPW; ++t) {
        if (pk[t] < 0) continue;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            double2 v = make_double2(acc[t][q][0], acc[t][q][1]);
            reinterpret_cast<double2*>(out + ((size_t)(e0 + t) * 2 + q) * 64)[lane] = v;
        }
    }
}

// G{K,M}[8 ti + r][8 tj + c] = sum over CTAs; one CTA of 128 threads per entry (both matrices)
__global__ void __launch_bounds__(128)
k_gram_sym2_reduce(const double* __restrict__ partial, int nparts, const __grid_constant__ GramPlan plan,
                   double* __restrict__ GK, double* __restrict__ GM, int64_t ldg) {
    const int e = blockIdx.x;
    const int q = threadIdx.x >> 6, idx = threadIdx.x & 63;      // matrix, element of the tile in lane order
    double s = 0.0;
    for (int c = 0; c < nparts; ++c) s += partial[((size_t)c * GS_MAX_ENTRIES + e) * 128 + q * 64 + idx];
    const int lane = idx >> 1, half = idx & 1;
    const int r = lane >> 2, col = 2 * (lane & 3) + half;
    double* G = q == 0 ? GK : GM;
    G[(int64_t)(8 * plan.ti[e] + r) * ldg + 8 * plan.tj[e] + col] = s;
}

int64_t gram_sym2_scratch_elems(int num_sms) { return (int64_t)num_sms * GS_MAX_ENTRIES * 128; }

// tiles[0..ntiles): ascending active tile columns (8 columns each) of the ld-wide blocks.  Writes the
// upper-triangle tiles (ti <= tj) of GK, GM (row-major, ldg); everything else is left untouched.
int gram_sym2(const double* S, const double* KS, const double* MS, int64_t ld, int64_t n, const int* tiles, int ntiles,
              double* GK, double* GM, int64_t ldg, double* partial, int num_sms, cudaStream_t stream) {
    DS_REQUIRE(S && KS && MS && GK && GM && partial && tiles, "gram_sym2: null argument");
    DS_REQUIRE(ld % 8 == 0 && ld <= 144 && ld > 0, "gram_sym2: ld=%lld must be a multiple of 8, <= 144", (long long)ld);
    DS_REQUIRE(ntiles >= 1 && ntiles <= 18, "gram_sym2: 1..18 active tiles");
    DS_REQUIRE(((uintptr_t)S % 16 == 0) && ((uintptr_t)KS % 16 == 0) && ((uintptr_t)MS % 16 == 0),
               "gram_sym2: blocks must be 16-byte aligned");
    DS_REQUIRE(n > 0, "gram_sym2: n must be positive");
    GramPlan plan;
    plan.n_entries = 0;
    for (int a = 0; a < ntiles; ++a)            // sorted by the A-side tile: a warp's consecutive entries share it
        for (int b = a; b < ntiles; ++b) {
            DS_REQUIRE(tiles[a] >= 0 && tiles[b] < ld / 8 && tiles[a] <= tiles[b], "gram_sym2: bad tile list");
            plan.ti[plan.n_entries] = (unsigned char)tiles[a];
            plan.tj[plan.n_entries] = (unsigned char)tiles[b];
            plan.n_entries++;
        }
    ProfScope prof(PROF_GRAM, stream);
    prof_account(PROF_GRAM, 3.0 * (double)n * ld * 8.0, 2.0 * 2.0 * (double)n * 64.0 * plan.n_entries);
    static bool attr = false;
    if (!attr) {
        DS_CUDA(cudaFuncSetAttribute(k_gram_sym2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GS_SMEM));
        attr = true;
    }
    const int64_t chunks = ceil_div(n, GS_ROWS);
    const int ctas = (int)(chunks < num_sms ? chunks : num_sms);
    k_gram_sym2<<<ctas, GS_THREADS, GS_SMEM, stream>>>(S, KS, MS, ld, (int)ld, n, plan, partial);
    DS_LAUNCH_CHECK();
    k_gram_sym2_reduce<<<plan.n_entries, 128, 0, stream>>>(partial, ctas, plan, GK, GM, ldg);
    DS_LAUNCH_CHECK();
    return DS_OK;
}

}  // namespace ds

using namespace ds;

extern "C" int64_t ds_gram_sym2_scratch_elems(void) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return -1;
    return gram_sym2_scratch_elems(sms);
}

extern "C" int ds_gram_sym2_f64(const double* S, const double* KS, const double* MS, int64_t ld, int64_t n,
                                const int* tiles_host, int ntiles, double* GK, double* GM, int64_t ldg, double* partial,
                                void* stream) {
    int dev = 0, sms = 0;
    DS_CUDA(cudaGetDevice(&dev));
    DS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    return gram_sym2(S, KS, MS, ld, n, tiles_host, ntiles, GK, GM, ldg, partial, sms, (cudaStream_t)stream);
}
