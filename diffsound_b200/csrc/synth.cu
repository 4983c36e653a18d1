// Modal synthesis: bank of damped sinusoids, forward and backward, FP32.
//
// Reference behaviour replaced (/root/reference/src/ddsp/oscillator.py:297-304, also :128-138, :160-171,
// :230-238): cumsum(d/sr), cumsum(f/sr) over a materialised (B, k, T) tensor, exp, sin, product, sum
// over modes.  cumsum of a constant is (t+1) c / sr, so
//     y[b,t] = sum_m a[b,m] s_m(t),   s_m(t) = exp(-d_m tau) sin(2 pi f_m tau),  tau = (t+1)/sr.
// Damping and damped frequency are per mode (shared by the batch) in every oscillator of the
// reference, so the basis s_m(t) is generated once per (mode, time tile) in shared memory by a
// complex phase recurrence z <- z w, w = exp((-d + 2 pi i f)/sr), re-anchored from an fp64 phase
// every SEG samples, and the batch is a register-tiled FP32 contraction against it.  Nothing of
// size B*k*T ever exists.
//
// Backward (gy = dL/dy):
//     gamp[b,m] = sum_t gy[b,t] s_m(t)
//     z[m,t]    = sum_b a[b,m] gy[b,t]
//     gdamp[m]  = sum_t -tau s_m(t) z[m,t],   gfreq[m] = sum_t 2 pi tau c_m(t) z[m,t]
// with c_m the cosine partner.  Partial sums over time tiles are reduced in a fixed order.
#include "common.cuh"
#include "../../include/diffsound_sm100.h"

namespace ds {

constexpr int SY_THREADS = 256;
constexpr int SY_BT = 128;   // time samples per tile
constexpr int SY_BB = 64;    // batch rows per tile
constexpr int SY_MK = 32;    // modes per chunk
constexpr int SY_SEG = 16;   // recurrence length between fp64 anchors

// z = exp(-d tau) (cos, sin)(2 pi f tau) at tau = (t+1)/sr, anchored in fp64
__device__ __forceinline__ void anchor(float d, float f, int64_t t, double inv_sr, float& re, float& im) {
    double tau = (double)(t + 1) * inv_sr;
    double ph = (double)f * tau;
    ph -= floor(ph);
    float s, c;
    sincospif((float)(2.0 * ph), &s, &c);
    float dec = (float)exp(-(double)d * tau);
    re = dec * c;
    im = dec * s;
}

__device__ __forceinline__ void rotor(float d, float f, double inv_sr, float& wr, float& wi) {
    double s, c;
    sincospi(2.0 * (double)f * inv_sr, &s, &c);
    double dec = exp(-(double)d * inv_sr);
    wr = (float)(dec * c);
    wi = (float)(dec * s);
}

// Fill S[mk][t] (and optionally C[mk][t]) for modes m0..m0+MK, times t0..t0+BT.
template <bool WITH_COS>
__device__ __forceinline__ void fill_basis(const float* __restrict__ damp, const float* __restrict__ freq, int k,
                                           int m0, int64_t t0, double inv_sr, float (*S)[SY_BT + 4],
                                           float (*Cc)[SY_BT + 4]) {
    // 32 modes x 8 segments of 16 samples = 256 threads
    const int mk = threadIdx.x >> 3, seg = threadIdx.x & 7;
    const int m = m0 + mk;
    float re = 0.f, im = 0.f, wr = 0.f, wi = 0.f;
    if (m < k) {
        float d = __ldg(damp + m), f = __ldg(freq + m);
        anchor(d, f, t0 + seg * SY_SEG, inv_sr, re, im);
        rotor(d, f, inv_sr, wr, wi);
    }
#pragma unroll
    for (int j = 0; j < SY_SEG; ++j) {
        S[mk][seg * SY_SEG + j] = im;
        if (WITH_COS) Cc[mk][seg * SY_SEG + j] = re;
        float nr = re * wr - im * wi;
        float ni = re * wi + im * wr;
        re = nr;
        im = ni;
    }
}

// ---------------------------------------------------------------------------
// forward: grid (ceil(T/BT), ceil(B/BB))
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(SY_THREADS)
k_synth_fwd(const float* __restrict__ amp, const float* __restrict__ damp, const float* __restrict__ freq, int64_t B,
            int k, int64_t T, double inv_sr, float* __restrict__ y) {
    __shared__ __align__(16) float S[SY_MK][SY_BT + 4];
    __shared__ __align__(16) float A[SY_MK][SY_BB + 4];
    const int64_t t0 = (int64_t)blockIdx.x * SY_BT;
    const int64_t b0 = (int64_t)blockIdx.y * SY_BB;
    const int tx = threadIdx.x & 15;   // 16 x 8 time samples
    const int ty = threadIdx.x >> 4;   // 16 x 4 batch rows
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int m0 = 0; m0 < k; m0 += SY_MK) {
        __syncthreads();
        fill_basis<false>(damp, freq, k, m0, t0, inv_sr, S, nullptr);
        for (int idx = threadIdx.x; idx < SY_MK * SY_BB; idx += SY_THREADS) {
            int bb = idx / SY_MK, mk = idx - bb * SY_MK;   // consecutive threads read consecutive modes
            int64_t b = b0 + bb;
            int m = m0 + mk;
            A[mk][bb] = (b < B && m < k) ? __ldg(amp + b * k + m) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int mk = 0; mk < SY_MK; ++mk) {
            float4 a4 = *reinterpret_cast<const float4*>(&A[mk][ty * 4]);
            float4 s0 = *reinterpret_cast<const float4*>(&S[mk][tx * 8]);
            float4 s1 = *reinterpret_cast<const float4*>(&S[mk][tx * 8 + 4]);
            float a[4] = {a4.x, a4.y, a4.z, a4.w};
            float s[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], s[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t b = b0 + ty * 4 + i;
        if (b >= B) continue;
        int64_t t = t0 + tx * 8;
        float* yp = y + b * T + t;
        if (t + 8 <= T && ((reinterpret_cast<uintptr_t>(yp) & 15) == 0)) {
            *reinterpret_cast<float4*>(yp) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            *reinterpret_cast<float4*>(yp + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (t + j < T) yp[j] = acc[i][j];
        }
    }
}

// ---------------------------------------------------------------------------
// backward 1: gamp partials.  grid (n_chunks over T, ceil(B/BB), ceil(k/MK)); each CTA owns a
// 64 x 32 (batch x mode) output tile and a contiguous run of time tiles.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(SY_THREADS)
k_synth_bwd_amp(const float* __restrict__ damp, const float* __restrict__ freq, const float* __restrict__ gy,
                int64_t B, int k, int64_t T, double inv_sr, int tiles_per_chunk, float* __restrict__ partial) {
    __shared__ __align__(16) float S[SY_MK][SY_BT + 4];
    constexpr int HT = SY_BT / 2;   // gy is staged half a time tile at a time (48 KB static limit)
    __shared__ __align__(16) float Gy[SY_BB][HT + 4];
    const int64_t b0 = (int64_t)blockIdx.y * SY_BB;
    const int m0 = blockIdx.z * SY_MK;
    const int tx = threadIdx.x & 7;    // 8 x 4 modes
    const int ty = threadIdx.x >> 3;   // 32 x 2 batch rows
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    const int64_t n_tiles = (T + SY_BT - 1) / SY_BT;
    const int64_t tile_lo = (int64_t)blockIdx.x * tiles_per_chunk;
    const int64_t tile_hi = min(n_tiles, tile_lo + tiles_per_chunk);
    for (int64_t tile = tile_lo; tile < tile_hi; ++tile) {
        const int64_t t0 = tile * SY_BT;
        __syncthreads();
        fill_basis<false>(damp, freq, k, m0, t0, inv_sr, S, nullptr);
        for (int h = 0; h < 2; ++h) {
            if (h) __syncthreads();
            for (int idx = threadIdx.x; idx < SY_BB * HT; idx += SY_THREADS) {
                int bb = idx / HT, tt = idx - bb * HT;
                int64_t b = b0 + bb, t = t0 + h * HT + tt;
                Gy[bb][tt] = (b < B && t < T) ? __ldg(gy + b * T + t) : 0.f;
            }
            __syncthreads();
#pragma unroll 4
            for (int tt = 0; tt < HT; tt += 4) {
                float4 g0 = *reinterpret_cast<const float4*>(&Gy[ty * 2][tt]);
                float4 g1 = *reinterpret_cast<const float4*>(&Gy[ty * 2 + 1][tt]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float4 s = *reinterpret_cast<const float4*>(&S[tx * 4 + j][h * HT + tt]);
                    acc[0][j] = fmaf(g0.x, s.x, fmaf(g0.y, s.y, fmaf(g0.z, s.z, fmaf(g0.w, s.w, acc[0][j]))));
                    acc[1][j] = fmaf(g1.x, s.x, fmaf(g1.y, s.y, fmaf(g1.z, s.z, fmaf(g1.w, s.w, acc[1][j]))));
                }
            }
        }
    }
    // partial[chunk][b][m]
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        int64_t b = b0 + ty * 2 + i;
        if (b >= B) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int m = m0 + tx * 4 + j;
            if (m < k) partial[((int64_t)blockIdx.x * B + b) * k + m] = acc[i][j];
        }
    }
}

// ---------------------------------------------------------------------------
// backward 2: z = A^T gy per (mode chunk, time tile), contracted at once with -tau s and 2 pi tau c.
// grid (ceil(T/BT), ceil(k/MK)); partial[tile][2][k]
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(SY_THREADS)
k_synth_bwd_df(const float* __restrict__ amp, const float* __restrict__ damp, const float* __restrict__ freq,
               const float* __restrict__ gy, int64_t B, int k, int64_t T, double inv_sr,
               float* __restrict__ partial) {
    // the (A, Gy) staging tiles of the batch loop and the (S, C) basis tiles of the epilogue share storage
    __shared__ __align__(16) float raw[2 * SY_MK * (SY_BT + 4)];
    __shared__ float red[2][SY_MK][17];
    float (*S)[SY_BT + 4] = reinterpret_cast<float (*)[SY_BT + 4]>(raw);
    float (*Cc)[SY_BT + 4] = reinterpret_cast<float (*)[SY_BT + 4]>(raw + SY_MK * (SY_BT + 4));
    float (*Gy)[SY_BT + 4] = reinterpret_cast<float (*)[SY_BT + 4]>(raw);                               // [bb][t]
    float (*A)[SY_MK + 4] = reinterpret_cast<float (*)[SY_MK + 4]>(raw + SY_MK * (SY_BT + 4));          // [bb][mk]
    const int64_t t0 = (int64_t)blockIdx.x * SY_BT;
    const int m0 = blockIdx.y * SY_MK;
    const int tx = threadIdx.x & 15;   // 16 x 8 time samples
    const int ty = threadIdx.x >> 4;   // 16 x 2 modes
    float z[2][8];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) z[i][j] = 0.f;
    for (int64_t b0 = 0; b0 < B; b0 += SY_MK) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < SY_MK * SY_MK; idx += SY_THREADS) {
            int bb = idx / SY_MK, mk = idx - bb * SY_MK;
            int64_t b = b0 + bb;
            int m = m0 + mk;
            A[bb][mk] = (b < B && m < k) ? __ldg(amp + b * k + m) : 0.f;
        }
        for (int idx = threadIdx.x; idx < SY_MK * SY_BT; idx += SY_THREADS) {
            int bb = idx / SY_BT, tt = idx - bb * SY_BT;
            int64_t b = b0 + bb, t = t0 + tt;
            Gy[bb][tt] = (b < B && t < T) ? __ldg(gy + b * T + t) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int bb = 0; bb < SY_MK; ++bb) {
            float a0 = A[bb][ty * 2], a1 = A[bb][ty * 2 + 1];
            float4 g0 = *reinterpret_cast<const float4*>(&Gy[bb][tx * 8]);
            float4 g1 = *reinterpret_cast<const float4*>(&Gy[bb][tx * 8 + 4]);
            float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                z[0][j] = fmaf(a0, g[j], z[0][j]);
                z[1][j] = fmaf(a1, g[j], z[1][j]);
            }
        }
    }
    __syncthreads();
    fill_basis<true>(damp, freq, k, m0, t0, inv_sr, S, Cc);
    __syncthreads();
    // contract over this thread's 8 samples, then over the 16 tx lanes
    const float two_pi = 6.283185307179586f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        int mk = ty * 2 + i;
        float gd = 0.f, gf = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int tt = tx * 8 + j;
            float tau = (float)((double)(t0 + tt + 1) * inv_sr);
            gd = fmaf(-tau * S[mk][tt], z[i][j], gd);
            gf = fmaf(two_pi * tau * Cc[mk][tt], z[i][j], gf);
        }
        red[0][mk][tx] = gd;
        red[1][mk][tx] = gf;
    }
    __syncthreads();
    if (threadIdx.x < 2 * SY_MK) {
        int which = threadIdx.x / SY_MK, mk = threadIdx.x - which * SY_MK;
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 16; ++q) s += red[which][mk][q];
        int m = m0 + mk;
        if (m < k) partial[((int64_t)blockIdx.x * 2 + which) * k + m] = s;
    }
}

// out[i] = sum_p partial[p][i]   (double accumulation, fixed order)
__global__ void k_synth_reduce(const float* __restrict__ partial, int64_t nparts, int64_t width,
                               float* __restrict__ out0, int64_t split, float* __restrict__ out1) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= width) return;
    double s = 0.0;
    for (int64_t p = 0; p < nparts; ++p) s += (double)partial[p * width + i];
    if (i < split) out0[i] = (float)s;
    else out1[i - split] = (float)s;
}

static int amp_chunks(int64_t T) {
    int64_t n_tiles = ceil_div(T, SY_BT);
    int64_t chunks = ceil_div(n_tiles, 32);   // up to 32 time tiles (4096 samples) per CTA
    return (int)(chunks < 1 ? 1 : chunks);
}

}  // namespace ds

using namespace ds;

extern "C" int64_t ds_synth_scratch_elems(int64_t B, int k, int64_t T) {
    int64_t a = (int64_t)amp_chunks(T) * B * k;
    int64_t d = ceil_div(T, SY_BT) * 2 * k;
    return a > d ? a : d;
}

extern "C" int ds_modal_synth_fwd(const float* amp, const float* damp, const float* freq, int64_t B, int k,
                                  int64_t T, double sr, float* y, float* scratch, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    (void)scratch;
    DS_REQUIRE(amp && damp && freq && y, "ds_modal_synth_fwd: null argument");
    DS_REQUIRE(B > 0 && k > 0 && T > 0 && sr > 0, "ds_modal_synth_fwd: bad sizes (B=%lld k=%d T=%lld)", (long long)B, k,
               (long long)T);
    dim3 grid((unsigned)ceil_div(T, SY_BT), (unsigned)ceil_div(B, SY_BB));
    ProfScope prof(PROF_SYNTH, stream);
    DS_REQUIRE(grid.y <= 65535, "ds_modal_synth_fwd: batch too large");
    k_synth_fwd<<<grid, SY_THREADS, 0, stream>>>(amp, damp, freq, B, k, T, 1.0 / sr, y);
    DS_LAUNCH_CHECK();
    return DS_OK;
}

extern "C" int ds_modal_synth_bwd(const float* amp, const float* damp, const float* freq, const float* gy, int64_t B,
                                  int k, int64_t T, double sr, float* gamp, float* gdamp, float* gfreq,
                                  float* scratch, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DS_REQUIRE(amp && damp && freq && gy && gamp && gdamp && gfreq && scratch, "ds_modal_synth_bwd: null argument");
    DS_REQUIRE(B > 0 && k > 0 && T > 0 && sr > 0, "ds_modal_synth_bwd: bad sizes");
    const int chunks = amp_chunks(T);
    ProfScope prof(PROF_SYNTH, stream);
    const int64_t n_tiles = ceil_div(T, SY_BT);
    const int tiles_per_chunk = (int)ceil_div(n_tiles, chunks);
    dim3 g1((unsigned)chunks, (unsigned)ceil_div(B, SY_BB), (unsigned)ceil_div(k, SY_MK));
    DS_REQUIRE(g1.y <= 65535 && g1.z <= 65535, "ds_modal_synth_bwd: batch or mode count too large");
    k_synth_bwd_amp<<<g1, SY_THREADS, 0, stream>>>(damp, freq, gy, B, k, T, 1.0 / sr, tiles_per_chunk, scratch);
    DS_LAUNCH_CHECK();
    k_synth_reduce<<<(unsigned)ceil_div(B * k, 256), 256, 0, stream>>>(scratch, chunks, B * k, gamp, B * k, nullptr);
    DS_LAUNCH_CHECK();
    dim3 g2((unsigned)n_tiles, (unsigned)ceil_div(k, SY_MK));
    k_synth_bwd_df<<<g2, SY_THREADS, 0, stream>>>(amp, damp, freq, gy, B, k, T, 1.0 / sr, scratch);
    DS_LAUNCH_CHECK();
    k_synth_reduce<<<(unsigned)ceil_div(2 * (int64_t)k, 256), 256, 0, stream>>>(scratch, n_tiles, 2 * (int64_t)k, gdamp, k,
                                                                                gfreq);
    DS_LAUNCH_CHECK();
    return DS_OK;
}
