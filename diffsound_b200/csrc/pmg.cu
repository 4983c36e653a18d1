// Coarse (linear, P1) level of the two-level p-multigrid preconditioner for quadratic tets.
//
// Reference context: the quadratic mesh is the linear mesh plus one mid-edge node per edge
// (/root/reference/src/diffelastic/mesh.py:116-160, local order [v0 m01 v1 m12 v2 m02 m03 m13 m23 v3]);
// the P1 space on the same tets is the subspace "mid-edge value = mean of the edge's end points", so
//   * prolongation P: corner -> itself, mid-edge -> 0.5 (end point a + end point b),
//   * the Galerkin coarse operator P^T K P is the P1 stiffness matrix, which the order-1
//     assembly kernel produces directly from the corner nodes (csrc/assemble.cu).
// This file does the integer work: corner numbering, the coarse tet list, the parent table of P and
// the fixed-order gather lists of P^T.  The reference has no counterpart (it factorises on the CPU).
#include "common.cuh"
#include "../../include/diffsound_sm100.h"
#include <cub/cub.cuh>

namespace ds {

__constant__ int c_corner_local[4] = {0, 2, 4, 9};
// mid-edge local index -> the two corner local indices of its edge
__constant__ int c_mid_local[6] = {1, 3, 5, 6, 7, 8};
__constant__ int c_mid_par[6][2] = {{0, 2}, {2, 4}, {4, 0}, {0, 9}, {2, 9}, {4, 9}};

__global__ void k_mark_corners(const int32_t* __restrict__ tets, int64_t T, int32_t* __restrict__ flag) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= 4 * T) return;
    const int64_t e = t >> 2;
    flag[tets[e * 10 + c_corner_local[t & 3]]] = 1;
}

// cid[i] = coarse id (rank among corner nodes, ascending fine id) or -1
__global__ void k_corner_ids(const int32_t* __restrict__ flag, const int32_t* __restrict__ scan, int64_t n_nodes,
                             int32_t* __restrict__ cid) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    cid[i] = flag[i] ? scan[i] : -1;
}

__global__ void k_coarse_nodes(const int32_t* __restrict__ cid, const float* __restrict__ verts, int64_t n_nodes,
                               float* __restrict__ cverts, int32_t* __restrict__ parents) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const int32_t c = cid[i];
    if (c >= 0) {
        cverts[3 * (int64_t)c + 0] = verts[3 * i + 0];
        cverts[3 * (int64_t)c + 1] = verts[3 * i + 1];
        cverts[3 * (int64_t)c + 2] = verts[3 * i + 2];
        parents[2 * i] = c;
        parents[2 * i + 1] = c;
    }
}

__global__ void k_coarse_tets(const int32_t* __restrict__ tets, int64_t T, const int32_t* __restrict__ cid,
                              int32_t* __restrict__ ctets, int32_t* __restrict__ parents) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= 10 * T) return;
    const int64_t e = t / 10;
    const int r = (int)(t - e * 10);
    const int32_t* te = tets + e * 10;
    if (r < 4) {
        ctets[4 * e + r] = cid[te[c_corner_local[r]]];
    } else {
        const int q = r - 4;
        const int32_t a = cid[te[c_mid_par[q][0]]], b = cid[te[c_mid_par[q][1]]];
        const int64_t mid = te[c_mid_local[q]];
        // every tet around the edge writes the same (min, max) pair
        parents[2 * mid] = min(a, b);
        parents[2 * mid + 1] = max(a, b);
    }
}

__global__ void k_iota2(int64_t n2, int32_t* __restrict__ vals) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n2) vals[t] = (int32_t)(t >> 1);
}

// rptr[I] = first position of key >= I in the sorted key list (I = 0..n_coarse)
__global__ void k_lower_bounds(const int32_t* __restrict__ keys, int64_t n2, int64_t n_coarse, int32_t* __restrict__ rptr) {
    const int64_t I = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (I > n_coarse) return;
    int64_t lo = 0, hi = n2;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < (int32_t)I) lo = mid + 1; else hi = mid;
    }
    rptr[I] = (int32_t)lo;
}

}  // namespace ds

using namespace ds;

extern "C" int ds_pmg_coarse_count(ds_workspace* ws, const int32_t* tets, int64_t T, int64_t n_nodes, int32_t* cid,
                                   int64_t* n_coarse_host, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    DS_REQUIRE(ws && tets && cid && n_coarse_host, "ds_pmg_coarse_count: null argument");
    DS_REQUIRE(T > 0 && n_nodes > 0 && n_nodes < (int64_t)1 << 30, "ds_pmg_coarse_count: bad mesh size");
    ProfScope prof(PROF_PATTERN, st);
    size_t scan_tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (int32_t*)nullptr, (int32_t*)nullptr, (int)n_nodes, st);
    DS_TRY(ws->arena.reserve((size_t)n_nodes * 8 + scan_tmp + 4096, st));
    int32_t* flag = ws->arena.take<int32_t>(n_nodes);
    int32_t* scan = ws->arena.take<int32_t>(n_nodes);
    void* tmp = ws->arena.take<char>(scan_tmp);
    DS_REQUIRE(tmp != nullptr, "ds_pmg_coarse_count: arena too small");
    DS_CUDA(cudaMemsetAsync(flag, 0, sizeof(int32_t) * n_nodes, st));
    k_mark_corners<<<(unsigned)ceil_div(4 * T, 256), 256, 0, st>>>(tets, T, flag);
    DS_LAUNCH_CHECK();
    DS_CUDA(cub::DeviceScan::ExclusiveSum(tmp, scan_tmp, flag, scan, (int)n_nodes, st));
    k_corner_ids<<<(unsigned)ceil_div(n_nodes, 256), 256, 0, st>>>(flag, scan, n_nodes, cid);
    DS_LAUNCH_CHECK();
    int32_t last[2] = {0, 0};
    DS_CUDA(cudaMemcpyAsync(&last[0], scan + (n_nodes - 1), 4, cudaMemcpyDeviceToHost, st));
    DS_CUDA(cudaMemcpyAsync(&last[1], flag + (n_nodes - 1), 4, cudaMemcpyDeviceToHost, st));
    DS_CUDA(cudaStreamSynchronize(st));
    *n_coarse_host = (int64_t)last[0] + last[1];
    return DS_OK;
}

extern "C" int ds_pmg_coarse_fill(ds_workspace* ws, const float* verts, const int32_t* tets, int64_t T,
                                  int64_t n_nodes, const int32_t* cid, int64_t n_coarse, int32_t* ctets,
                                  float* cverts, int32_t* parents, int32_t* rptr, int32_t* rlist, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    DS_REQUIRE(ws && verts && tets && cid && ctets && cverts && parents && rptr && rlist,
               "ds_pmg_coarse_fill: null argument");
    DS_REQUIRE(((uintptr_t)parents & 7) == 0, "ds_pmg_coarse_fill: parents must be 8-byte aligned");
    ProfScope prof(PROF_PATTERN, st);
    const int64_t n2 = 2 * n_nodes;
    int end_bit = 1;
    while (end_bit < 31 && ((int64_t)1 << end_bit) < n_coarse) ++end_bit;
    size_t sort_tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (int32_t*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr,
                                    (int32_t*)nullptr, (int)n2, 0, end_bit, st);
    DS_TRY(ws->arena.reserve((size_t)n2 * 8 + sort_tmp + 4096, st));
    int32_t* keys_out = ws->arena.take<int32_t>(n2);
    int32_t* vals_in = ws->arena.take<int32_t>(n2);
    void* tmp = ws->arena.take<char>(sort_tmp);
    DS_REQUIRE(tmp != nullptr, "ds_pmg_coarse_fill: arena too small");
    DS_CUDA(cudaMemsetAsync(parents, 0, sizeof(int32_t) * n2, st));
    k_coarse_nodes<<<(unsigned)ceil_div(n_nodes, 256), 256, 0, st>>>(cid, verts, n_nodes, cverts, parents);
    DS_LAUNCH_CHECK();
    k_coarse_tets<<<(unsigned)ceil_div(10 * T, 256), 256, 0, st>>>(tets, T, cid, ctets, parents);
    DS_LAUNCH_CHECK();
    k_iota2<<<(unsigned)ceil_div(n2, 256), 256, 0, st>>>(n2, vals_in);
    DS_LAUNCH_CHECK();
    // stable sort by parent: each coarse node's gather list is ascending in the fine node id
    DS_CUDA(cub::DeviceRadixSort::SortPairs(tmp, sort_tmp, parents, keys_out, vals_in, rlist, (int)n2, 0, end_bit, st));
    k_lower_bounds<<<(unsigned)ceil_div(n_coarse + 1, 256), 256, 0, st>>>(keys_out, n2, n_coarse, rptr);
    DS_LAUNCH_CHECK();
    return DS_OK;
}
