// DDSP filtered noise: time-varying FIR-filtered uniform noise, forward and backward.
//
// Reference behaviour replaced: FilteredNoise.forward (/root/reference/src/ddsp/filtered_noise.py:20-67), used by
// GTDampedOscillator.forward as `signal + noise * noise_rate` (/root/reference/src/ddsp/oscillator.py:206,226,243;
// material_real_train.py:118 trains it for 2001 epochs with noise_rate = 2e-4).  The reference goes through five
// FFTs per call (irfft of the zero-phase response, rfft of the windowed impulse response and of the zero-padded noise,
// irfft of the product) and an overlap-add written as a conv_transpose1d with an identity kernel.  The filters are
// short (2C - 1 = 129 taps) and the frames shorter (L = 64), so here everything stays in the time domain:
//
//   x_k   = 2 sigmoid(c_k)^2.3 + 1e-6                                   (ddsp/utils.py:6-9), k < C
//   h0[t] = (x_0 + 2 sum_{k>=1} x_k cos(2 pi k t / N)) / N,  N = 2C - 1  (irfft of a real spectrum, odd N)
//   h[t]  = hann_N[t] * h0[(t - (C - 1)) mod N]                         (roll + periodic Hann window)
//   fr[j] = gain * sum_i noise[i] h[j - i],  0 <= j < L + N - 1          (linear convolution = the padded FFT product)
//   y[t]  = sum_f fr_f[t - L f]                                          (overlap-add, cropped to T)
//
// One CTA per output segment of L samples gathers the three frames that overlap it (deterministic, no atomics);
// the backward pass is one CTA per frame: d h = correlation of the upstream gradient with the frame's noise,
// d x = cosine transform of (window * d h), d c through the modified sigmoid.
#include "common.cuh"
#include "../../include/diffsound_sm100.h"

namespace ds {

constexpr int FN_MAXC = 129;               // filter_coeff_length (reference default 65)
constexpr int FN_MAXN = 2 * FN_MAXC - 1;   // taps
constexpr int FN_MAXL = 256;               // frame_length (reference default 64)
constexpr int FN_THREADS = 256;

__device__ __forceinline__ float modified_sigmoid(float c) {
    const float s = 1.f / (1.f + expf(-c));
    return 2.f * powf(s, 2.3f) + 1e-6f;
}

// h[0..N) of one frame into shared memory; xs: scratch [C], cs: cos table [N]
__device__ void frame_ir(const float* __restrict__ coeff, int C, int N, const float* cs, float* xs, float* h) {
    for (int k = threadIdx.x; k < C; k += blockDim.x) xs[k] = modified_sigmoid(coeff[k]);
    __syncthreads();
    for (int t = threadIdx.x; t < N; t += blockDim.x) {
        int t0 = t - (C - 1);
        if (t0 < 0) t0 += N;
        float acc = 0.f;
        for (int k = 1; k < C; ++k) acc = fmaf(xs[k], cs[(int)(((long long)k * t0) % N)], acc);
        const float h0 = (xs[0] + 2.f * acc) / (float)N;
        const float w = 0.5f - 0.5f * cs[t];                  // periodic Hann of length N: 0.5 - 0.5 cos(2 pi t / N)
        h[t] = w * h0;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(FN_THREADS)
k_filtered_noise_fwd(const float* __restrict__ coeff, const float* __restrict__ noise, int F, int C, int L, int64_t T,
                     float gain, float* __restrict__ y) {
    __shared__ float cs[FN_MAXN], xs[FN_MAXC], h[FN_MAXN], nz[FN_MAXL], acc[FN_MAXL];
    const int N = 2 * C - 1, seg = blockIdx.x, b = blockIdx.y;
    for (int t = threadIdx.x; t < N; t += blockDim.x) cs[t] = cospif(2.f * (float)t / (float)N);
    for (int j = threadIdx.x; j < L; j += blockDim.x) acc[j] = 0.f;
    __syncthreads();
    const int span = (L + N - 2) / L;                          // frames before `seg` that still reach into it
    for (int f = seg - span; f <= seg; ++f) {                  // ascending frame order: a fixed summation order
        if (f < 0 || f >= F) continue;
        frame_ir(coeff + ((size_t)b * F + f) * C, C, N, cs, xs, h);
        for (int i = threadIdx.x; i < L; i += blockDim.x) nz[i] = noise[((size_t)b * F + f) * L + i];
        __syncthreads();
        const int off = (seg - f) * L;                          // position of this segment inside the frame's output
        for (int j = threadIdx.x; j < L; j += blockDim.x) {
            float s = 0.f;
            const int jj = off + j;
            for (int i = 0; i < L; ++i) {
                const int tau = jj - i;
                if (tau >= 0 && tau < N) s = fmaf(nz[i], h[tau], s);
            }
            acc[j] += gain * s;
        }
        __syncthreads();
    }
    for (int j = threadIdx.x; j < L; j += blockDim.x) {
        const int64_t t = (int64_t)seg * L + j;
        if (t < T) y[(size_t)b * T + t] = acc[j];
    }
}

__global__ void __launch_bounds__(FN_THREADS)
k_filtered_noise_bwd(const float* __restrict__ coeff, const float* __restrict__ noise, const float* __restrict__ gy,
                     int F, int C, int L, int64_t T, float gain, float* __restrict__ gcoeff) {
    __shared__ float cs[FN_MAXN], gh[FN_MAXN], nz[FN_MAXL], g[FN_MAXL + FN_MAXN];
    const int N = 2 * C - 1, f = blockIdx.x, b = blockIdx.y;
    for (int t = threadIdx.x; t < N; t += blockDim.x) cs[t] = cospif(2.f * (float)t / (float)N);
    for (int i = threadIdx.x; i < L; i += blockDim.x) nz[i] = noise[((size_t)b * F + f) * L + i];
    for (int j = threadIdx.x; j < L + N - 1; j += blockDim.x) {
        const int64_t t = (int64_t)f * L + j;
        g[j] = t < T ? gain * gy[(size_t)b * T + t] : 0.f;
    }
    __syncthreads();
    // d h[tau] = sum_i noise[i] g[tau + i], times the window
    for (int tau = threadIdx.x; tau < N; tau += blockDim.x) {
        float s = 0.f;
        for (int i = 0; i < L; ++i) s = fmaf(nz[i], g[tau + i], s);
        gh[tau] = s * (0.5f - 0.5f * cs[tau]);
    }
    __syncthreads();
    // d x_k = (k ? 2 : 1) / N sum_tau gh[tau] cos(2 pi k (tau - (C - 1)) / N);  d c_k = d x_k * 4.6 s^2.3 (1 - s)
    for (int k = threadIdx.x; k < C; k += blockDim.x) {
        float s = 0.f;
        for (int tau = 0; tau < N; ++tau) {
            int t0 = tau - (C - 1);
            if (t0 < 0) t0 += N;
            s = fmaf(gh[tau], cs[(int)(((long long)k * t0) % N)], s);
        }
        s *= (k ? 2.f : 1.f) / (float)N;
        const float c = coeff[((size_t)b * F + f) * C + k];
        const float sg = 1.f / (1.f + expf(-c));
        gcoeff[((size_t)b * F + f) * C + k] = s * 4.6f * powf(sg, 2.3f) * (1.f - sg);
    }
}

}  // namespace ds

using namespace ds;

static int fn_check(const char* who, int64_t B, int F, int C, int L, int64_t T) {
    DS_REQUIRE(B > 0 && B <= 65535 && F > 0 && T > 0, "%s: bad sizes (B=%lld F=%d T=%lld)", who, (long long)B, F, (long long)T);
    DS_REQUIRE(C >= 2 && C <= FN_MAXC && L >= 1 && L <= FN_MAXL, "%s: filter_coeff_length=%d (<= %d), frame_length=%d (<= %d)",
               who, C, FN_MAXC, L, FN_MAXL);
    DS_REQUIRE((int64_t)F * L >= T, "%s: %d frames of %d samples do not cover %lld samples", who, F, L, (long long)T);
    return DS_OK;
}

extern "C" int ds_filtered_noise_fwd(const float* coeff, const float* noise, int64_t B, int F, int C, int L, int64_t T,
                                     double gain, float* y, void* stream) {
    DS_REQUIRE(coeff && noise && y, "ds_filtered_noise_fwd: null argument");
    DS_TRY(fn_check("ds_filtered_noise_fwd", B, F, C, L, T));
    ProfScope prof(PROF_SYNTH, (cudaStream_t)stream);
    const int segs = (int)ceil_div(T, L);
    k_filtered_noise_fwd<<<dim3(segs, (unsigned)B), FN_THREADS, 0, (cudaStream_t)stream>>>(coeff, noise, F, C, L, T,
                                                                                            (float)gain, y);
    DS_LAUNCH_CHECK();
    return DS_OK;
}

extern "C" int ds_filtered_noise_bwd(const float* coeff, const float* noise, const float* gy, int64_t B, int F, int C,
                                     int L, int64_t T, double gain, float* gcoeff, void* stream) {
    DS_REQUIRE(coeff && noise && gy && gcoeff, "ds_filtered_noise_bwd: null argument");
    DS_TRY(fn_check("ds_filtered_noise_bwd", B, F, C, L, T));
    ProfScope prof(PROF_SYNTH, (cudaStream_t)stream);
    k_filtered_noise_bwd<<<dim3(F, (unsigned)B), FN_THREADS, 0, (cudaStream_t)stream>>>(coeff, noise, gy, F, C, L, T,
                                                                                         (float)gain, gcoeff);
    DS_LAUNCH_CHECK();
    return DS_OK;
}
