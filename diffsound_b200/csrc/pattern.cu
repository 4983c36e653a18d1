// Sparsity pattern of K and M: node-level block CSR + contributor lists.
//
// Reference behaviour replaced: the pattern that sparse_coo_tensor(...).coalesce()
// re-derives by sort-and-reduce for every 20 000-Gauss-point batch
// (/root/reference/src/diffelastic/diff_model.py:214-220, 305-312).  Here it is built once
// per topology: one stable radix sort of the T*npe^2 (node_i, node_j) element
// entries; the sorted order doubles as the per-slot contributor list that lets
// assembly run owner-computes (no atomics).
#include "common.cuh"
#include "../../include/diffsound_sm100.h"
#include <cub/cub.cuh>
#include <mutex>

namespace ds {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int Arena::reserve(size_t bytes, cudaStream_t s) {
    used = 0;
    if (have_stream && last_stream != s) {          // the scratch changes hands between streams: order the new user behind the old
        if (!order_ev) DS_CUDA(cudaEventCreateWithFlags(&order_ev, cudaEventDisableTiming));
        DS_CUDA(cudaEventRecord(order_ev, last_stream));
        DS_CUDA(cudaStreamWaitEvent(s, order_ev, 0));
    }
    last_stream = s;
    have_stream = true;
    if (bytes <= cap) return DS_OK;
    if (base) {
        DS_CUDA(cudaStreamSynchronize(s));
        DS_CUDA(cudaFree(base));
        base = nullptr;
        cap = 0;
    }
    size_t want = bytes + (bytes >> 3) + (1 << 20);
    cudaError_t e = cudaMalloc(&base, want);
    if (e != cudaSuccess) {
        set_error("workspace cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        base = nullptr;
        return DS_ERR_NOMEM;
    }
    cap = want;
    return DS_OK;
}

void Arena::release() {
    if (base) cudaFree(base);
    if (order_ev) cudaEventDestroy(order_ev);
    base = nullptr;
    order_ev = nullptr;
    have_stream = false;
    cap = used = 0;
}

__global__ void k_gen_pairs(const int32_t* __restrict__ tets, int64_t n_pairs, int npe,
                            int64_t n_nodes, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= n_pairs) return;
    int npe2 = npe * npe;
    int64_t e = q / npe2;
    int r = (int)(q - e * npe2);
    int a = r / npe, b = r - a * npe;
    uint64_t i = (uint64_t)tets[e * npe + a];
    uint64_t j = (uint64_t)tets[e * npe + b];
    keys[q] = i * (uint64_t)n_nodes + j;
    vals[q] = (uint32_t)q;
}

__global__ void k_mark_heads(const uint64_t* __restrict__ keys, int64_t n, uint32_t* __restrict__ flag) {
    int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= n) return;
    flag[q] = (q == 0 || keys[q] != keys[q - 1]) ? 1u : 0u;
}

__global__ void k_fill(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                       const uint32_t* __restrict__ slot_incl, int64_t n_pairs, int64_t n_nodes,
                       int32_t* __restrict__ brow, int32_t* __restrict__ bcol,
                       int32_t* __restrict__ contrib_ptr, int32_t* __restrict__ contrib,
                       int32_t* __restrict__ slot) {
    int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= n_pairs) return;
    uint32_t s = slot_incl[q] - 1;
    uint32_t v = vals[q];
    contrib[q] = (int32_t)v;
    if (slot) slot[v] = (int32_t)s;
    uint64_t key = keys[q];
    int64_t i = (int64_t)(key / (uint64_t)n_nodes);
    bool head = (q == 0) || (slot_incl[q - 1] != slot_incl[q]);
    if (head) {
        contrib_ptr[s] = (int32_t)q;
        bcol[s] = (int32_t)(key - (uint64_t)i * (uint64_t)n_nodes);
        int64_t ip = (q == 0) ? -1 : (int64_t)(keys[q - 1] / (uint64_t)n_nodes);
        for (int64_t r = ip + 1; r <= i; ++r) brow[r] = (int32_t)s;
    }
    if (q == n_pairs - 1) {
        uint32_t nnzb = slot_incl[q];
        contrib_ptr[nnzb] = (int32_t)n_pairs;
        for (int64_t r = i + 1; r <= n_nodes; ++r) brow[r] = (int32_t)nnzb;
    }
}

__global__ void k_expand_csr(const int32_t* __restrict__ brow, const int32_t* __restrict__ bcol,
                             int64_t n_nodes, int64_t* __restrict__ crow, int64_t* __restrict__ col) {
    // one warp per node row
    int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (w >= n_nodes) return;
    int64_t b0 = brow[w], b1 = brow[w + 1];
    int64_t deg = b1 - b0;
    if (lane < 3) crow[3 * w + lane] = 9 * b0 + lane * 3 * deg;
    if (w == n_nodes - 1 && lane == 0) crow[3 * n_nodes] = 9 * b1;
    for (int64_t t = lane; t < 3 * deg; t += 32) {
        int64_t p = t / 3;
        int d = (int)(t - 3 * p);
        int64_t c = 3 * (int64_t)bcol[b0 + p] + d;
        col[9 * b0 + t] = c;
        col[9 * b0 + 3 * deg + t] = c;
        col[9 * b0 + 6 * deg + t] = c;
    }
}

}  // namespace ds

using namespace ds;

extern "C" int ds_version(void) { return 100; }
extern "C" const char* ds_last_error(void) { return ds::g_err; }

extern "C" int ds_workspace_create(ds_workspace** ws) {
    DS_REQUIRE(ws != nullptr, "ds_workspace_create: null out pointer");
    ds_workspace* w = new ds_workspace();
    int dev = 0;
    DS_CUDA(cudaGetDevice(&dev));
    DS_CUDA(cudaDeviceGetAttribute(&w->num_sms, cudaDevAttrMultiProcessorCount, dev));
    *ws = w;
    return DS_OK;
}

extern "C" int ds_workspace_destroy(ds_workspace* ws) {
    if (!ws) return DS_OK;
    if (ws->child) ds_workspace_destroy(ws->child);
    ws->arena.release();
    delete ws;
    return DS_OK;
}

extern "C" int64_t ds_workspace_bytes(const ds_workspace* ws) {
    return ws ? (int64_t)ws->arena.cap + (ws->child ? (int64_t)ws->child->arena.cap : 0) : 0;
}

extern "C" int ds_pattern_count(ds_workspace* ws, const int32_t* tets, int64_t T, int npe,
                                int64_t n_nodes, int64_t* nnzb_host, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DS_REQUIRE(ws && tets && nnzb_host, "ds_pattern_count: null argument");
    DS_REQUIRE(npe == 4 || npe == 10, "ds_pattern_count: npe must be 4 or 10 (got %d)", npe);
    DS_REQUIRE(T > 0 && n_nodes > 0, "ds_pattern_count: empty mesh (T=%lld, n_nodes=%lld)", (long long)T,
               (long long)n_nodes);
    int64_t n_pairs = T * npe * npe;
    ProfScope prof(PROF_PATTERN, stream);
    DS_REQUIRE(n_pairs < (int64_t)2147483647, "ds_pattern_count: T*npe^2 = %lld exceeds int32", (long long)n_pairs);
    DS_REQUIRE(9 * n_pairs < (int64_t)1 << 40, "too large");
    int end_bit = 1;
    {
        unsigned __int128 mx = (unsigned __int128)n_nodes * (unsigned __int128)n_nodes;
        while (end_bit < 64 && (((unsigned __int128)1) << end_bit) < mx) ++end_bit;
    }
    size_t sort_tmp = 0, scan_tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (uint64_t*)nullptr, (uint64_t*)nullptr,
                                    (uint32_t*)nullptr, (uint32_t*)nullptr, (int)n_pairs, 0, end_bit, stream);
    cub::DeviceScan::InclusiveSum(nullptr, scan_tmp, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)n_pairs, stream);
    size_t tmp = sort_tmp > scan_tmp ? sort_tmp : scan_tmp;
    size_t total = (size_t)n_pairs * (8 + 8 + 4 + 4 + 4 + 4) + tmp + 16 * 256;
    DS_TRY(ws->arena.reserve(total, stream));
    uint64_t* keys_in = ws->arena.take<uint64_t>(n_pairs);
    uint64_t* keys_out = ws->arena.take<uint64_t>(n_pairs);
    uint32_t* vals_in = ws->arena.take<uint32_t>(n_pairs);
    uint32_t* vals_out = ws->arena.take<uint32_t>(n_pairs);
    uint32_t* flag = ws->arena.take<uint32_t>(n_pairs);
    uint32_t* slot_incl = ws->arena.take<uint32_t>(n_pairs);
    void* cub_tmp = ws->arena.take<char>(tmp);
    DS_REQUIRE(cub_tmp != nullptr, "ds_pattern_count: arena too small");
    int threads = 256;
    int blocks = (int)ceil_div(n_pairs, threads);
    k_gen_pairs<<<blocks, threads, 0, stream>>>(tets, n_pairs, npe, n_nodes, keys_in, vals_in);
    DS_LAUNCH_CHECK();
    DS_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, sort_tmp, keys_in, keys_out, vals_in, vals_out,
                                            (int)n_pairs, 0, end_bit, stream));
    k_mark_heads<<<blocks, threads, 0, stream>>>(keys_out, n_pairs, flag);
    DS_LAUNCH_CHECK();
    DS_CUDA(cub::DeviceScan::InclusiveSum(cub_tmp, scan_tmp, flag, slot_incl, (int)n_pairs, stream));
    uint32_t nnzb = 0;
    DS_CUDA(cudaMemcpyAsync(&nnzb, slot_incl + (n_pairs - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    DS_CUDA(cudaStreamSynchronize(stream));
    ws->sorted_keys = keys_out;
    ws->sorted_vals = vals_out;
    ws->slot_of_sorted = slot_incl;
    ws->n_pairs = n_pairs;
    ws->nnzb = nnzb;
    *nnzb_host = nnzb;
    return DS_OK;
}

extern "C" int ds_pattern_fill(ds_workspace* ws, int64_t n_nodes, int32_t* brow, int32_t* bcol,
                               int32_t* contrib_ptr, int32_t* contrib, int32_t* slot, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DS_REQUIRE(ws && ws->sorted_keys, "ds_pattern_fill: call ds_pattern_count first");
    DS_REQUIRE(brow && bcol && contrib_ptr && contrib, "ds_pattern_fill: null output");
    int threads = 256;
    int blocks = (int)ceil_div(ws->n_pairs, threads);
    k_fill<<<blocks, threads, 0, stream>>>(ws->sorted_keys, ws->sorted_vals, ws->slot_of_sorted, ws->n_pairs,
                                           n_nodes, brow, bcol, contrib_ptr, contrib, slot);
    DS_LAUNCH_CHECK();
    ws->sorted_keys = nullptr;  // arena may be reused after this
    return DS_OK;
}

extern "C" int ds_pattern_expand_csr(const int32_t* brow, const int32_t* bcol, int64_t n_nodes, int64_t nnzb,
                                     int64_t* crow, int64_t* col, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DS_REQUIRE(brow && bcol && crow && col, "ds_pattern_expand_csr: null argument");
    (void)nnzb;
    int threads = 256;
    int64_t blocks = ceil_div(n_nodes * 32, threads);
    k_expand_csr<<<(unsigned)blocks, threads, 0, stream>>>(brow, bcol, n_nodes, crow, col);
    DS_LAUNCH_CHECK();
    return DS_OK;
}
