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

// one-sample rotor w = exp((-d + 2 pi i f) / sr).  FP32 transcendental on an argument formed in fp64: the
// recurrence is re-anchored from an fp64 phase every SY_SEG samples, so the rotor's 1e-7 relative error grows to at
// most ~2e-6 before it is discarded (audio tolerance 1e-4), and the fp64 sincospi / exp it replaces cost as much
// as a third of the contraction in the forward kernel.
__device__ __forceinline__ void rotor(float d, float f, double inv_sr, float& wr, float& wi) {
    float s, c;
    sincospif((float)(2.0 * (double)f * inv_sr), &s, &c);
    const float dec = expf(-(float)((double)d * inv_sr));
    wr = dec * c;
    wi = dec * s;
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
// forward: grid (ceil(T/BT), ceil(B/SF_BB)).  y tile = 128 batch rows x 128 samples per CTA, 8 x 8 per thread
// (rows 4 ty + {0..3} and 64 + 4 ty + {0..3}, samples 4 tx + {0..3} and 64 + 4 tx + {0..3}: every shared-memory
// read is one conflict-free 128-bit load).  Per mode a thread does 4 LDS.128 and 32 packed FFMA2
// (fma.rn.f32x2 with the amplitude broadcast) = 64 FMAs: the FP32 pipe, not shared-memory bandwidth, is
// the limit (the 4 x 8 tile with scalar FFMA ran at 96 % of the shared-memory pipe and 41 % of the FMA pipe).
// ---------------------------------------------------------------------------
constexpr int SF_BB = 128;   // batch rows per forward tile

__device__ __forceinline__ void ffma2_bcast(unsigned long long& acc, float a, unsigned long long s2) {
    unsigned long long aa;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(aa), "l"(s2));
}

__global__ void __launch_bounds__(SY_THREADS, 2)
k_synth_fwd(const float* __restrict__ amp, const float* __restrict__ damp, const float* __restrict__ freq, int64_t B,
            int k, int64_t T, double inv_sr, float* __restrict__ y) {
    __shared__ __align__(16) float S[SY_MK][SY_BT + 4];
    __shared__ __align__(16) float A[SY_MK][SF_BB + 4];
    const int64_t t0 = (int64_t)blockIdx.x * SY_BT;
    const int64_t b0 = (int64_t)blockIdx.y * SF_BB;
    const int tx = threadIdx.x & 15;   // samples 4 tx + {0..3}, 64 + 4 tx + {0..3}
    const int ty = threadIdx.x >> 4;   // rows    4 ty + {0..3}, 64 + 4 ty + {0..3}
    unsigned long long acc[8][4];      // [row][sample pair]
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0ull;
    for (int m0 = 0; m0 < k; m0 += SY_MK) {
        __syncthreads();
        fill_basis<false>(damp, freq, k, m0, t0, inv_sr, S, nullptr);
        for (int idx = threadIdx.x; idx < SY_MK * SF_BB; idx += SY_THREADS) {
            int bb = idx / SY_MK, mk = idx - bb * SY_MK;   // consecutive threads read consecutive modes
            int64_t b = b0 + bb;
            int m = m0 + mk;
            A[mk][bb] = (b < B && m < k) ? __ldg(amp + b * k + m) : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int mk = 0; mk < SY_MK; ++mk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&A[mk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&A[mk][64 + ty * 4]);
            const ulonglong2 s0 = *reinterpret_cast<const ulonglong2*>(&S[mk][tx * 4]);
            const ulonglong2 s1 = *reinterpret_cast<const ulonglong2*>(&S[mk][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                ffma2_bcast(acc[i][0], a[i], s0.x);
                ffma2_bcast(acc[i][1], a[i], s0.y);
                ffma2_bcast(acc[i][2], a[i], s1.x);
                ffma2_bcast(acc[i][3], a[i], s1.y);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t b = b0 + (i < 4 ? 4 * ty + i : 64 + 4 * ty + (i - 4));
        if (b >= B) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t t = t0 + 64 * h + 4 * tx;
            float v[4];
            asm("mov.b64 {%0, %1}, %2;" : "=f"(v[0]), "=f"(v[1]) : "l"(acc[i][2 * h]));
            asm("mov.b64 {%0, %1}, %2;" : "=f"(v[2]), "=f"(v[3]) : "l"(acc[i][2 * h + 1]));
            float* yp = y + b * T + t;
            if (t + 4 <= T && ((reinterpret_cast<uintptr_t>(yp) & 15) == 0)) {
                *reinterpret_cast<float4*>(yp) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (t + j < T) yp[j] = v[j];
            }
        }
    }
}

// ---------------------------------------------------------------------------
// backward 1: gamp[b,m] = sum_t gy[b,t] s_m(t), partial sums per time chunk.
// grid (n_chunks over T, ceil(B/128), ceil(k/128)); each CTA owns a 128 x 128 (batch x mode) output tile and a
// contiguous run of 32-sample slabs; 8 x 8 outputs per thread (rows 4 ty + {0..3}, 64 + 4 ty + {0..3}; modes
// 4 tx + {0..3}, 64 + 4 tx + {0..3}).  Both operands are staged TRANSPOSED (sample-major), so per sample a thread
// does 4 conflict-free LDS.128 and 32 FFMA2 -- the same FP32-pipe-bound inner loop as the forward kernel.
// ---------------------------------------------------------------------------
constexpr int SB_TT = 32;    // samples per staged slab
constexpr int SB_W = 128;    // tile width (batch rows, modes or samples)

__global__ void __launch_bounds__(SY_THREADS, 2)
k_synth_bwd_amp(const float* __restrict__ damp, const float* __restrict__ freq, const float* __restrict__ gy,
                int64_t B, int k, int64_t T, double inv_sr, int slabs_per_chunk, float* __restrict__ partial) {
    __shared__ __align__(16) float ST[SB_TT][SB_W + 4];    // [sample][mode]
    __shared__ __align__(16) float GT[SB_TT][SB_W + 4];    // [sample][batch row]
    const int64_t b0 = (int64_t)blockIdx.y * SB_W;
    const int m0 = blockIdx.z * SB_W;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    unsigned long long acc[8][4];      // [row][mode pair]
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0ull;
    const int64_t n_slabs = (T + SB_TT - 1) / SB_TT;
    const int64_t slab_lo = (int64_t)blockIdx.x * slabs_per_chunk;
    const int64_t slab_hi = min(n_slabs, slab_lo + slabs_per_chunk);
    // basis generator of this thread: one mode, one 16-sample segment of every slab
    const int gm = threadIdx.x & (SB_W - 1), gseg = threadIdx.x >> 7;
    float gd = 0.f, gf = 0.f, wr = 0.f, wi = 0.f;
    const bool gvalid = m0 + gm < k;
    if (gvalid) {
        gd = __ldg(damp + m0 + gm);
        gf = __ldg(freq + m0 + gm);
        rotor(gd, gf, inv_sr, wr, wi);
    }
    for (int64_t slab = slab_lo; slab < slab_hi; ++slab) {
        const int64_t t0 = slab * SB_TT;
        __syncthreads();
        {
            float re = 0.f, im = 0.f;
            if (gvalid) anchor(gd, gf, t0 + gseg * SY_SEG, inv_sr, re, im);
#pragma unroll
            for (int j = 0; j < SY_SEG; ++j) {
                ST[gseg * SY_SEG + j][gm] = im;
                const float nr = re * wr - im * wi;
                im = re * wi + im * wr;
                re = nr;
            }
        }
#pragma unroll
        for (int r = 0; r < (SB_W * SB_TT / 4) / SY_THREADS; ++r) {
            const int idx = threadIdx.x + r * SY_THREADS;
            const int bb = idx & (SB_W - 1), q = idx >> 7;          // 4 samples 4 q .. 4 q + 3 of row bb
            const int64_t b = b0 + bb, t = t0 + 4 * q;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (b < B) {
                const float* gp = gy + b * T + t;
                if (t + 4 <= T && (reinterpret_cast<uintptr_t>(gp) & 15) == 0) {
                    const float4 w = __ldg(reinterpret_cast<const float4*>(gp));
                    v[0] = w.x; v[1] = w.y; v[2] = w.z; v[3] = w.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (t + e < T) v[e] = __ldg(gp + e);
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) GT[4 * q + e][bb] = v[e];
        }
        __syncthreads();
#pragma unroll 4
        for (int tt = 0; tt < SB_TT; ++tt) {
            const float4 g0 = *reinterpret_cast<const float4*>(&GT[tt][ty * 4]);
            const float4 g1 = *reinterpret_cast<const float4*>(&GT[tt][64 + ty * 4]);
            const ulonglong2 s0 = *reinterpret_cast<const ulonglong2*>(&ST[tt][tx * 4]);
            const ulonglong2 s1 = *reinterpret_cast<const ulonglong2*>(&ST[tt][64 + tx * 4]);
            const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                ffma2_bcast(acc[i][0], g[i], s0.x);
                ffma2_bcast(acc[i][1], g[i], s0.y);
                ffma2_bcast(acc[i][2], g[i], s1.x);
                ffma2_bcast(acc[i][3], g[i], s1.y);
            }
        }
    }
    // partial[chunk][b][m]
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t b = b0 + (i < 4 ? 4 * ty + i : 64 + 4 * ty + (i - 4));
        if (b >= B) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int m = m0 + 64 * h + 4 * tx;
            float v[4];
            asm("mov.b64 {%0, %1}, %2;" : "=f"(v[0]), "=f"(v[1]) : "l"(acc[i][2 * h]));
            asm("mov.b64 {%0, %1}, %2;" : "=f"(v[2]), "=f"(v[3]) : "l"(acc[i][2 * h + 1]));
            float* pp = partial + ((int64_t)blockIdx.x * B + b) * k + m;
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (m + e < k) pp[e] = v[e];
        }
    }
}

// ---------------------------------------------------------------------------
// backward 2: z[m,t] = sum_b a[b,m] gy[b,t] per (128 modes x 128 samples) tile, 8 x 8 per thread with the same
// FFMA2 inner loop (reduction over the batch in slabs of 32 rows), contracted in registers with -tau s_m(t) and
// 2 pi tau c_m(t) (basis regenerated per thread from fp64 anchors), reduced over the 16 sample lanes by shuffles.
// grid (ceil(T/128), ceil(k/128)); partial[tile][2][k]
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(SY_THREADS, 2)
k_synth_bwd_df(const float* __restrict__ amp, const float* __restrict__ damp, const float* __restrict__ freq,
               const float* __restrict__ gy, int64_t B, int k, int64_t T, double inv_sr,
               float* __restrict__ partial) {
    __shared__ __align__(16) float A[SB_TT][SB_W + 4];     // [batch row][mode]
    __shared__ __align__(16) float G[SB_TT][SB_W + 4];     // [batch row][sample]
    const int64_t t0 = (int64_t)blockIdx.x * SB_W;
    const int m0 = blockIdx.y * SB_W;
    const int tx = threadIdx.x & 15;   // samples 4 tx + {0..3}, 64 + 4 tx + {0..3}
    const int ty = threadIdx.x >> 4;   // modes   4 ty + {0..3}, 64 + 4 ty + {0..3}
    unsigned long long z[8][4];        // [mode][sample pair]
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) z[i][j] = 0ull;
    for (int64_t b0 = 0; b0 < B; b0 += SB_TT) {
        __syncthreads();
#pragma unroll
        for (int r = 0; r < (SB_W * SB_TT) / SY_THREADS; ++r) {
            const int idx = threadIdx.x + r * SY_THREADS;
            const int c = idx & (SB_W - 1), bb = idx >> 7;
            const int64_t b = b0 + bb;
            A[bb][c] = (b < B && m0 + c < k) ? __ldg(amp + b * k + m0 + c) : 0.f;
            G[bb][c] = (b < B && t0 + c < T) ? __ldg(gy + b * T + t0 + c) : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int bb = 0; bb < SB_TT; ++bb) {
            const float4 a0 = *reinterpret_cast<const float4*>(&A[bb][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&A[bb][64 + ty * 4]);
            const ulonglong2 g0 = *reinterpret_cast<const ulonglong2*>(&G[bb][tx * 4]);
            const ulonglong2 g1 = *reinterpret_cast<const ulonglong2*>(&G[bb][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                ffma2_bcast(z[i][0], a[i], g0.x);
                ffma2_bcast(z[i][1], a[i], g0.y);
                ffma2_bcast(z[i][2], a[i], g1.x);
                ffma2_bcast(z[i][3], a[i], g1.y);
            }
        }
    }
    const float two_pi = 6.283185307179586f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? 4 * ty + i : 64 + 4 * ty + (i - 4));
        float gd = 0.f, gf = 0.f;
        if (m < k) {
            const float d = __ldg(damp + m), f = __ldg(freq + m);
            float wr, wi;
            rotor(d, f, inv_sr, wr, wi);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int64_t t = t0 + 64 * h + 4 * tx;
                float re, im;
                anchor(d, f, t, inv_sr, re, im);
                float zz[4];
                asm("mov.b64 {%0, %1}, %2;" : "=f"(zz[0]), "=f"(zz[1]) : "l"(z[i][2 * h]));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(zz[2]), "=f"(zz[3]) : "l"(z[i][2 * h + 1]));
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (t + e < T) {
                        const float tau = (float)((double)(t + e + 1) * inv_sr);
                        gd = fmaf(-tau * im, zz[e], gd);
                        gf = fmaf(two_pi * tau * re, zz[e], gf);
                    }
                    const float nr = re * wr - im * wi;
                    im = re * wi + im * wr;
                    re = nr;
                }
            }
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            gd += __shfl_xor_sync(0xffffffffu, gd, o);
            gf += __shfl_xor_sync(0xffffffffu, gf, o);
        }
        if (tx == 0 && m < k) {
            partial[((int64_t)blockIdx.x * 2 + 0) * k + m] = gd;
            partial[((int64_t)blockIdx.x * 2 + 1) * k + m] = gf;
        }
    }
}

// ---------------------------------------------------------------------------
// Causal force FIR applied to the rendered audio (oscillator.py:305-309: conv1d with the flipped force, groups =
// audio_num, padding F-1, cropped to T):   out[b,t] = sum_{i<F} force[b,i] x[b,t-i]      (reverse = 0)
// and its adjoint w.r.t. x (backward):      out[b,t] = sum_{i<F} force[b,i] x[b,t+i]      (reverse = 1).
// One CTA per (1024-sample tile, audio): the force and the tile (+ F-1 halo samples) are staged in shared memory,
// each thread produces 4 consecutive samples.
// ---------------------------------------------------------------------------
constexpr int FIR_TILE = 1024;
constexpr int FIR_MAXF = 2048;

__global__ void __launch_bounds__(256)
k_force_fir(const float* __restrict__ x, const float* __restrict__ force, int64_t T, int F, int reverse,
            float* __restrict__ out) {
    extern __shared__ float fir_sh[];          // fs[F] | xs[FIR_TILE + F - 1]
    float* fs = fir_sh;
    float* xs = fir_sh + F;
    const int64_t b = blockIdx.y;
    const int64_t t0 = (int64_t)blockIdx.x * FIR_TILE;
    const float* xb = x + b * T;
    for (int i = threadIdx.x; i < F; i += blockDim.x) fs[i] = __ldg(force + b * F + i);
    // forward needs x[t0 - (F-1) .. t0 + TILE), reverse x[t0 .. t0 + TILE + F - 1)
    const int64_t base = reverse ? t0 : t0 - (F - 1);
    for (int i = threadIdx.x; i < FIR_TILE + F - 1; i += blockDim.x) {
        const int64_t t = base + i;
        xs[i] = (t >= 0 && t < T) ? __ldg(xb + t) : 0.f;
    }
    __syncthreads();
    const int l = threadIdx.x * 4;             // local sample of this thread's first output
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (!reverse) {
        // out[t0 + l + e] = sum_i fs[i] xs[(l + e) + (F-1) - i]
        for (int i = 0; i < F; ++i) {
            const float f = fs[i];
            const float* xp = xs + l + (F - 1) - i;
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[e] = fmaf(f, xp[e], acc[e]);
        }
    } else {
        for (int i = 0; i < F; ++i) {
            const float f = fs[i];
            const float* xp = xs + l + i;
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[e] = fmaf(f, xp[e], acc[e]);
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e)
        if (t0 + l + e < T) out[b * T + t0 + l + e] = acc[e];
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
    int64_t n_slabs = ceil_div(T, SB_TT);
    int64_t chunks = ceil_div(n_slabs, 75);   // up to 75 slabs (2400 samples) per CTA: 37 chunks at T = 88 200
    return (int)(chunks < 1 ? 1 : chunks);
}

}  // namespace ds

using namespace ds;

extern "C" int64_t ds_synth_scratch_elems(int64_t B, int k, int64_t T) {
    int64_t a = (int64_t)amp_chunks(T) * B * k;
    int64_t d = ceil_div(T, SB_W) * 2 * k;
    return a > d ? a : d;
}

extern "C" int ds_modal_synth_fwd(const float* amp, const float* damp, const float* freq, int64_t B, int k,
                                  int64_t T, double sr, float* y, float* scratch, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    (void)scratch;
    DS_REQUIRE(amp && damp && freq && y, "ds_modal_synth_fwd: null argument");
    DS_REQUIRE(B > 0 && k > 0 && T > 0 && sr > 0, "ds_modal_synth_fwd: bad sizes (B=%lld k=%d T=%lld)", (long long)B, k,
               (long long)T);
    dim3 grid((unsigned)ceil_div(T, SY_BT), (unsigned)ceil_div(B, SF_BB));
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
    const int64_t n_tiles = ceil_div(T, SB_W);
    const int slabs_per_chunk = (int)ceil_div(ceil_div(T, SB_TT), chunks);
    dim3 g1((unsigned)chunks, (unsigned)ceil_div(B, SB_W), (unsigned)ceil_div(k, SB_W));
    DS_REQUIRE(g1.y <= 65535 && g1.z <= 65535, "ds_modal_synth_bwd: batch or mode count too large");
    k_synth_bwd_amp<<<g1, SY_THREADS, 0, stream>>>(damp, freq, gy, B, k, T, 1.0 / sr, slabs_per_chunk, scratch);
    DS_LAUNCH_CHECK();
    k_synth_reduce<<<(unsigned)ceil_div(B * k, 256), 256, 0, stream>>>(scratch, chunks, B * k, gamp, B * k, nullptr);
    DS_LAUNCH_CHECK();
    dim3 g2((unsigned)n_tiles, (unsigned)ceil_div(k, SB_W));
    k_synth_bwd_df<<<g2, SY_THREADS, 0, stream>>>(amp, damp, freq, gy, B, k, T, 1.0 / sr, scratch);
    DS_LAUNCH_CHECK();
    k_synth_reduce<<<(unsigned)ceil_div(2 * (int64_t)k, 256), 256, 0, stream>>>(scratch, n_tiles, 2 * (int64_t)k, gdamp, k,
                                                                                gfreq);
    DS_LAUNCH_CHECK();
    return DS_OK;
}

extern "C" int ds_force_fir(const float* x, const float* force, int64_t B, int64_t T, int F, int reverse, float* out,
                            void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DS_REQUIRE(x && force && out, "ds_force_fir: null argument");
    DS_REQUIRE(B > 0 && T > 0 && F >= 1 && F <= FIR_MAXF, "ds_force_fir: bad sizes (B=%lld T=%lld F=%d, F <= %d)",
               (long long)B, (long long)T, F, FIR_MAXF);
    DS_REQUIRE(B <= 65535, "ds_force_fir: batch too large");
    DS_REQUIRE(x != out, "ds_force_fir: in-place operation is not supported");
    ProfScope prof(PROF_SYNTH, stream);
    dim3 grid((unsigned)ceil_div(T, FIR_TILE), (unsigned)B);
    const size_t smem = (size_t)(2 * F + FIR_TILE) * sizeof(float);
    k_force_fir<<<grid, 256, smem, stream>>>(x, force, T, F, reverse ? 1 : 0, out);
    DS_LAUNCH_CHECK();
    return DS_OK;
}
