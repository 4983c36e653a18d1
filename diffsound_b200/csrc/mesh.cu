// Node numbering of the promoted (quadratic) mesh: unique rows of an (N, 3) fp32 coordinate array in
// ascending lexicographic order, with the inverse map and the first original index of every group.
//
// Reference behaviour replaced (file:line under /root/reference/src):
//   diffelastic/mesh.py:162-179  TetMesh.remove_duplicate_vertices: torch.unique(vertices, dim=0,
//                                return_inverse=True) + scatter(min) of the original indices
//   diffelastic/mesh.py:101-160  to_high_order (the caller: V + 6T candidate nodes, 1.2 M rows at 200k tets)
// The numbering is part of the bit-exact pattern contract: node ids = rank of the
// coordinate triple in (x, y, z) lexicographic order of the fp32 values.
//
// Design: three stable LSD radix-sort passes (z, then y, then x) over order-preserving 32-bit keys of the
// floats (-0.0 canonicalised to +0.0 so it compares equal, like the float comparison of torch.unique),
// carrying the original index as the value; group heads by key comparison with the predecessor; an
// inclusive scan ranks the groups.  Stability makes the first element of a group its smallest original
// index, which is the representative scatter(min) picks.
#include "common.cuh"
#include "../../include/diffsound_sm100.h"
#include <cub/cub.cuh>

namespace ds {

__device__ __forceinline__ uint32_t float_key(float f) {
    f += 0.0f;                                     // -0.0 -> +0.0
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// keys[i] = key of coordinate `axis` of row idx[i] (idx == nullptr: identity, and vals[i] = i is written)
__global__ void k_axis_keys(const float* __restrict__ rows, int64_t N, int axis, const uint32_t* __restrict__ idx,
                            uint32_t* __restrict__ keys, uint32_t* __restrict__ iota) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int64_t r = idx ? (int64_t)idx[i] : i;
    keys[i] = float_key(rows[3 * r + axis]);
    if (iota) iota[i] = (uint32_t)i;
}

__global__ void k_row_heads(const float* __restrict__ rows, int64_t N, const uint32_t* __restrict__ idx,
                            uint32_t* __restrict__ head) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    uint32_t h = 1;
    if (i > 0) {
        const float* a = rows + 3 * (int64_t)idx[i];
        const float* b = rows + 3 * (int64_t)idx[i - 1];
        h = (float_key(a[0]) != float_key(b[0])) | (float_key(a[1]) != float_key(b[1])) |
            (float_key(a[2]) != float_key(b[2]));
    }
    head[i] = h;
}

__global__ void k_unique_fill(int64_t N, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ rank,
                              int64_t* __restrict__ inverse, int64_t* __restrict__ first) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    const uint32_t r = rank[i] - 1;
    inverse[idx[i]] = (int64_t)r;
    if (i == 0 || rank[i - 1] != rank[i]) first[r] = (int64_t)idx[i];
}

}  // namespace ds

using namespace ds;

extern "C" int ds_unique_rows3_count(ds_workspace* ws, const float* rows, int64_t N, int64_t* n_unique_host,
                                     void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    DS_REQUIRE(ws && rows && n_unique_host, "ds_unique_rows3_count: null argument");
    DS_REQUIRE(N > 0 && N < (int64_t)0x7fffffff, "ds_unique_rows3_count: N=%lld out of range", (long long)N);
    ProfScope prof(PROF_PATTERN, st);
    size_t tmp_sort = 0, tmp_scan = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)N, 0, 32, st);
    cub::DeviceScan::InclusiveSum(nullptr, tmp_scan, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)N, st);
    const size_t tmp = tmp_sort > tmp_scan ? tmp_sort : tmp_scan;
    DS_TRY(ws->arena.reserve(5 * ((size_t)N * 4 + 512) + tmp + 1024, st));
    Arena& a = ws->arena;
    uint32_t* k0 = a.take<uint32_t>(N);
    uint32_t* k1 = a.take<uint32_t>(N);
    uint32_t* v0 = a.take<uint32_t>(N);
    uint32_t* v1 = a.take<uint32_t>(N);
    uint32_t* rank = a.take<uint32_t>(N);
    void* scratch = a.take<char>(tmp);
    DS_REQUIRE(scratch != nullptr, "ds_unique_rows3_count: workspace arena exhausted");
    const unsigned blocks = (unsigned)ceil_div(N, 256);
    uint32_t *vin = v0, *vout = v1;
    for (int pass = 0; pass < 3; ++pass) {
        const int axis = 2 - pass;
        k_axis_keys<<<blocks, 256, 0, st>>>(rows, N, axis, pass == 0 ? nullptr : vin, k0, pass == 0 ? vin : nullptr);
        DS_LAUNCH_CHECK();
        size_t t = tmp;
        DS_CUDA(cub::DeviceRadixSort::SortPairs(scratch, t, k0, k1, vin, vout, (int)N, 0, 32, st));
        count_launch();
        uint32_t* s = vin; vin = vout; vout = s;
    }
    // vin: original indices in lexicographic row order
    k_row_heads<<<blocks, 256, 0, st>>>(rows, N, vin, k0);
    DS_LAUNCH_CHECK();
    size_t t = tmp;
    DS_CUDA(cub::DeviceScan::InclusiveSum(scratch, t, k0, rank, (int)N, st));
    count_launch();
    uint32_t n_u = 0;
    DS_CUDA(cudaMemcpyAsync(&n_u, rank + (N - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    DS_CUDA(cudaStreamSynchronize(st));
    *n_unique_host = (int64_t)n_u;
    ws->sorted_vals = vin;
    ws->slot_of_sorted = rank;
    ws->n_pairs = N;
    return DS_OK;
}

extern "C" int ds_unique_rows3_fill(ds_workspace* ws, int64_t* inverse, int64_t* first, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    DS_REQUIRE(ws && ws->sorted_vals && ws->slot_of_sorted, "ds_unique_rows3_fill: call ds_unique_rows3_count first");
    DS_REQUIRE(inverse && first, "ds_unique_rows3_fill: null output");
    const int64_t N = ws->n_pairs;
    k_unique_fill<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(N, ws->sorted_vals, ws->slot_of_sorted, inverse, first);
    DS_LAUNCH_CHECK();
    ws->sorted_vals = nullptr;
    ws->slot_of_sorted = nullptr;
    return DS_OK;
}
