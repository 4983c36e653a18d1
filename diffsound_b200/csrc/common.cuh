// Shared helpers for libdiffsound_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>

#define DS_OK 0
#define DS_ERR_CUDA -1
#define DS_ERR_ARG -2
#define DS_ERR_NOMEM -3
#define DS_ERR_NUMERIC -4

namespace ds {

void set_error(const char* fmt, ...);

inline int check_cuda(cudaError_t e, const char* what, const char* file, int line) {
    if (e == cudaSuccess) return DS_OK;
    set_error("%s failed at %s:%d: %s", what, file, line, cudaGetErrorString(e));
    return DS_ERR_CUDA;
}

#define DS_CUDA(call)                                                          \
    do {                                                                       \
        int _rc = ::ds::check_cuda((call), #call, __FILE__, __LINE__);         \
        if (_rc != DS_OK) return _rc;                                          \
    } while (0)

#define DS_LAUNCH_CHECK() DS_CUDA(cudaGetLastError())

#define DS_REQUIRE(cond, ...)                                                  \
    do {                                                                       \
        if (!(cond)) {                                                         \
            ::ds::set_error(__VA_ARGS__);                                      \
            return DS_ERR_ARG;                                                 \
        }                                                                      \
    } while (0)

#define DS_TRY(expr)                                                           \
    do {                                                                       \
        int _rc = (expr);                                                      \
        if (_rc != DS_OK) return _rc;                                          \
    } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Growable device scratch arena owned by a ds_workspace (the only allocation
// the library performs; everything else is caller-owned torch memory).
struct Arena {
    char* base = nullptr;
    size_t cap = 0;
    size_t used = 0;
    int reserve(size_t bytes, cudaStream_t s);
    void reset() { used = 0; }
    template <typename T>
    T* take(size_t count) {
        size_t off = (used + 255) & ~size_t(255);
        size_t need = off + count * sizeof(T);
        if (need > cap) return nullptr;
        used = need;
        return reinterpret_cast<T*>(base + off);
    }
    void release();
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace ds

struct ds_workspace {
    ds::Arena arena;
    // pattern-build state carried between ds_pattern_count and ds_pattern_fill
    uint64_t* sorted_keys = nullptr;
    uint32_t* sorted_vals = nullptr;
    uint32_t* slot_of_sorted = nullptr;
    int64_t n_pairs = 0;
    int64_t nnzb = 0;
    int num_sms = 0;
};
