// Eigenvalue derivatives: d(lambda)/d(vertices) and the material quadratic forms.
//
// Reference behaviour replaced (file:line under /root/reference/src):
//   diffelastic/diff_model.py:390-399  get_vals(): lambda + diag(U^T K U) - lambda diag(U^T M U), whose
//       autograd backward walks sparse-mm -> coalesce -> bmm -> torch.inverse -> vertices;
//   diffelastic/diff_model.py:314-328, 371-388 + deform.py:70-87, 149-165  the fp32 matrix-free
//       K(theta) U (gather -> F -> P -> scatter-add) behind get_undamped_freqs().
//
// Shape gradient.  For weights g_i the scalar E = sum_i g_i (u_i^T K u_i - lam_i u_i^T M u_i) splits per
// element.  With W = U_e diag(g) U_e^T (the weighted Gram of the element's rows of U, DMMA) and the
// same contraction table as assembly,
//     Q_lm = sum_ab ctab[a][b][l][m] W_ab              (3x3, 16 of them)
//     e(G) = sum_lm mu tr(Q_lm) G_l.G_m + mu G_m^T Q_lm G_l + lam G_l^T Q_lm G_m
//     E_K  = |det A| e(G),     E_M = -|det6V| sum_ab mtab[a][b] tr(W'_ab),  W' weighted by g_i lam_i
// e is a quadratic form in the 4x3 matrix G = dL/dxi A^-1, so dE/dG is linear in G and the chain
// rule to the four corner positions is closed form.  One warp per element; per-corner results go to
// tet_grad and are summed per node by a gather kernel (deterministic, no atomics).
//
// Material forms.  lane = mode.  Linear tets have a constant displacement gradient F; for quadratic
// tets F is linear in the barycentric coordinates, F(L) = sum_m L_m F_m with F_m its value at corner
// m, so the exact integral of a quadratic form phi is |det A|/120 (phi(sum_m F_m) + sum_m phi(F_m)).
#include "common.cuh"
#include "../../include/diffsound_sm100.h"
#include "kernels.cuh"
#include <cub/cub.cuh>

namespace ds {

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// nodes whose shape function depends on barycentric coordinate l (quadratic tet, reference local order)
__device__ __constant__ int c_L2[4][4] = {{0, 1, 5, 6}, {2, 1, 3, 7}, {4, 3, 5, 8}, {9, 6, 7, 8}};

struct TetGeom {
    double G[4][3];   // rows of dL/dxi A^-1
    double detK;      // |det A| (A from fp32 differences, as mesh.py:90-98)
    double detM;      // |6V| from fp64 corner coordinates (diff_model.py:272-289)
    double sgnM;      // sign of the signed 6V
    double dD[4][3];  // d(6V signed)/d(corner)
};

template <bool WITH_DD>
__device__ __forceinline__ void tet_geometry(const float* __restrict__ verts, const int32_t c[4], TetGeom& g) {
    float v[4][3];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        v[p][0] = __ldg(verts + 3 * (int64_t)c[p] + 0);
        v[p][1] = __ldg(verts + 3 * (int64_t)c[p] + 1);
        v[p][2] = __ldg(verts + 3 * (int64_t)c[p] + 2);
    }
    double A[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        A[r][0] = (double)__fsub_rn(v[0][r], v[3][r]);
        A[r][1] = (double)__fsub_rn(v[1][r], v[3][r]);
        A[r][2] = (double)__fsub_rn(v[2][r], v[3][r]);
    }
    double c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1];
    double c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2];
    double c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
    double det = A[0][0] * c00 + A[0][1] * c01 + A[0][2] * c02;
    double id = 1.0 / det;
    g.G[0][0] = c00 * id;
    g.G[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) * id;
    g.G[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) * id;
    g.G[1][0] = c01 * id;
    g.G[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * id;
    g.G[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) * id;
    g.G[2][0] = c02 * id;
    g.G[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) * id;
    g.G[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * id;
#pragma unroll
    for (int d = 0; d < 3; ++d) g.G[3][d] = -(g.G[0][d] + g.G[1][d] + g.G[2][d]);
    g.detK = fabs(det);
    double e1[3], e2[3], e3[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        e1[r] = (double)v[1][r] - (double)v[0][r];
        e2[r] = (double)v[2][r] - (double)v[0][r];
        e3[r] = (double)v[3][r] - (double)v[0][r];
    }
    double n1[3] = {e2[1] * e3[2] - e2[2] * e3[1], e2[2] * e3[0] - e2[0] * e3[2], e2[0] * e3[1] - e2[1] * e3[0]};
    double D = e1[0] * n1[0] + e1[1] * n1[1] + e1[2] * n1[2];
    g.detM = fabs(D);
    g.sgnM = D < 0.0 ? -1.0 : 1.0;
    if (WITH_DD) {
        double n2[3] = {e3[1] * e1[2] - e3[2] * e1[1], e3[2] * e1[0] - e3[0] * e1[2], e3[0] * e1[1] - e3[1] * e1[0]};
        double n3[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            g.dD[1][r] = n1[r];
            g.dD[2][r] = n2[r];
            g.dD[3][r] = n3[r];
            g.dD[0][r] = -(n1[r] + n2[r] + n3[r]);
        }
    }
}

__device__ __forceinline__ void load_corners(const int32_t* __restrict__ t, int order, int32_t c[4]) {
    if (order == 1) {
        c[0] = __ldg(t + 0); c[1] = __ldg(t + 1); c[2] = __ldg(t + 2); c[3] = __ldg(t + 3);
    } else {
        c[0] = __ldg(t + 0); c[1] = __ldg(t + 2); c[2] = __ldg(t + 4); c[3] = __ldg(t + 9);
    }
}

// ---------------------------------------------------------------------------
// shape gradient
// ---------------------------------------------------------------------------
constexpr int GS_WARPS = 4;
constexpr int GS_WLD = 33;   // leading dimension of the per-warp W tile in shared memory

template <int ORDER>
struct GradSmem {
    static constexpr int NPE = ORDER == 1 ? 4 : 10;
    static constexpr int ROWS = 3 * NPE;                 // 12 | 30
    static constexpr int RT = (ROWS + 7) / 8;            // 2 | 4 row tiles
    static constexpr int RP = RT * 8;                    // padded rows 16 | 32
    double ctab[NPE * NPE * 16];
    double mhat[RP * RP];                                // (mtab (x) I3), zero padded
    double W[GS_WARPS][RP * GS_WLD];
    double Q[GS_WARPS][192];                             // 144 Q entries, 12 dE/dG, 12 G (at 160), 12 dD (at 172)
};

template <int ORDER>
__global__ void __launch_bounds__(GS_WARPS * 32)
k_eigval_grad_shape(const float* __restrict__ verts, const int32_t* __restrict__ tets, int64_t T, double mu,
                    double lam, const double* __restrict__ ctab_g, const double* __restrict__ mtab_g,
                    const double* __restrict__ U, int64_t ldu, int k, const double* __restrict__ lamv,
                    const double* __restrict__ gv, double* __restrict__ tet_grad) {
    using SM = GradSmem<ORDER>;
    constexpr int NPE = SM::NPE, ROWS = SM::ROWS, RT = SM::RT, RP = SM::RP;
    constexpr int NL = ORDER == 1 ? 1 : 4;
    constexpr int NTILE = RT * (RT + 1) / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SM& sm = *reinterpret_cast<SM*>(smem_raw);
    double* s_w = reinterpret_cast<double*>(smem_raw + sizeof(SM));   // [2][kpad]: g, g*lam
    const int kpad = (k + 15) & ~15;
    for (int t = threadIdx.x; t < NPE * NPE * 16; t += blockDim.x) sm.ctab[t] = ctab_g[t];
    for (int t = threadIdx.x; t < RP * RP; t += blockDim.x) {
        int r = t / RP, c = t - r * RP;
        sm.mhat[t] = (r < ROWS && c < ROWS && (r % 3) == (c % 3)) ? mtab_g[(r / 3) * NPE + (c / 3)] : 0.0;
    }
    for (int t = threadIdx.x; t < kpad; t += blockDim.x) {
        double g = t < k ? gv[t] : 0.0;
        s_w[t] = g;
        s_w[kpad + t] = t < k ? g * lamv[t] : 0.0;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mm = lane >> 2, kk = lane & 3;
    double* Ws = sm.W[warp];
    double* Qs = sm.Q[warp];
    const int64_t wstride = (int64_t)gridDim.x * GS_WARPS;
    for (int64_t e = (int64_t)blockIdx.x * GS_WARPS + warp; e < T; e += wstride) {
        const int32_t* tp = tets + e * NPE;
        int32_t cn[4];
        load_corners(tp, ORDER, cn);
        TetGeom geo;
        tet_geometry<true>(verts, cn, geo);
        if (lane < 12) {   // lane-indexed reads of G and dD go through shared memory (no local-memory arrays)
            const int p = lane / 3, j = lane - 3 * p;
            double gv_ = 0.0, dv_ = 0.0;
#pragma unroll
            for (int pp = 0; pp < 4; ++pp)
#pragma unroll
                for (int jj = 0; jj < 3; ++jj)
                    if (pp == p && jj == j) { gv_ = geo.G[pp][jj]; dv_ = geo.dD[pp][jj]; }
            Qs[160 + lane] = gv_;
            Qs[172 + lane] = dv_;
        }
        const double* Gs = Qs + 160;
        // ---- weighted Grams of the element rows of U on the FP64 tensor pipe
        const double* rowp[RT];
#pragma unroll
        for (int ti = 0; ti < RT; ++ti) {
            int r = 8 * ti + mm;
            rowp[ti] = nullptr;
            if (r < ROWS) {
                int64_t node = __ldg(tp + r / 3);
                rowp[ti] = U + (3 * node + (r % 3)) * ldu;
            }
        }
        double accK[NTILE][2], accM[NTILE][2];
#pragma unroll
        for (int t = 0; t < NTILE; ++t) accK[t][0] = accK[t][1] = accM[t][0] = accM[t][1] = 0.0;
        for (int k0 = 0; k0 < kpad; k0 += 16) {
            const int kb = k0 + 4 * kk;   // this lane's four consecutive modes
            double u[RT][4];
#pragma unroll
            for (int ti = 0; ti < RT; ++ti) {
#pragma unroll
                for (int j = 0; j < 4; ++j) u[ti][j] = (rowp[ti] && kb + j < k) ? __ldg(rowp[ti] + kb + j) : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double wk = s_w[kb + j], wm = s_w[kpad + kb + j];
                double bK[RT], bM[RT];
#pragma unroll
                for (int ti = 0; ti < RT; ++ti) {
                    bK[ti] = u[ti][j] * wk;
                    bM[ti] = u[ti][j] * wm;
                }
                int t = 0;
#pragma unroll
                for (int ti = 0; ti < RT; ++ti)
#pragma unroll
                    for (int tj = ti; tj < RT; ++tj, ++t) {
                        dmma884(accK[t][0], accK[t][1], u[ti][j], bK[tj]);
                        dmma884(accM[t][0], accM[t][1], u[ti][j], bM[tj]);
                    }
            }
        }
        double s_mass = 0.0;
        {
            int t = 0;
#pragma unroll
            for (int ti = 0; ti < RT; ++ti)
#pragma unroll
                for (int tj = ti; tj < RT; ++tj, ++t) {
                    int r = 8 * ti + mm, c = 8 * tj + 2 * kk;
                    Ws[r * GS_WLD + c] = accK[t][0];
                    Ws[r * GS_WLD + c + 1] = accK[t][1];
                    if (ti != tj) {
                        Ws[c * GS_WLD + r] = accK[t][0];
                        Ws[(c + 1) * GS_WLD + r] = accK[t][1];
                    }
                    double f = ti == tj ? 1.0 : 2.0;
                    s_mass += f * (sm.mhat[r * RP + c] * accM[t][0] + sm.mhat[r * RP + c + 1] * accM[t][1]);
                }
        }
        s_mass = warp_sum(s_mass);
        __syncwarp();
        // ---- Q_lm = sum_ab ctab[a][b][l][m] W_ab
        for (int o = lane; o < 144; o += 32) {
            int lmi = o / 9, cd = o - lmi * 9;
            int l = lmi >> 2, m = lmi & 3, c = cd / 3, d = cd - 3 * c;
            double acc = 0.0;
#pragma unroll
            for (int ai = 0; ai < NL; ++ai) {
                int a = ORDER == 1 ? l : c_L2[l][ai];
#pragma unroll
                for (int bi = 0; bi < NL; ++bi) {
                    int b = ORDER == 1 ? m : c_L2[m][bi];
                    acc = fma(sm.ctab[((a * NPE + b) * 4 + l) * 4 + m], Ws[(3 * a + c) * GS_WLD + 3 * b + d], acc);
                }
            }
            Qs[o] = acc;
        }
        __syncwarp();
        // ---- dE/dG (lanes 0..11 = (p, j)) and e = 1/2 G : dE/dG
        double dG = 0.0;
        if (lane < 12) {
            int p = lane / 3, j = lane - 3 * p;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const double* Qpm = Qs + (p * 4 + m) * 9;
                const double* Qmp = Qs + (m * 4 + p) * 9;
                double tpm = Qpm[0] + Qpm[4] + Qpm[8];
                double s1 = Qmp[j * 3 + 0] * geo.G[m][0] + Qmp[j * 3 + 1] * geo.G[m][1] + Qmp[j * 3 + 2] * geo.G[m][2];
                double s2 = Qpm[j * 3 + 0] * geo.G[m][0] + Qpm[j * 3 + 1] * geo.G[m][1] + Qpm[j * 3 + 2] * geo.G[m][2];
                dG += 2.0 * mu * (tpm * Gs[3 * m + j] + s1) + 2.0 * lam * s2;
            }
        }
        double ee = 0.0;
        if (lane < 12) {
            int p = lane / 3, j = lane - 3 * p;
            ee = 0.5 * Gs[lane] * dG;
            Qs[144 + lane] = dG;
        }
        ee = warp_sum(ee);
        __syncwarp();
        // ---- chain rule to the four corners (lanes 0..11 = (corner p, coordinate r))
        if (lane < 12) {
            int p = lane / 3, r = lane - 3 * p;
            double B[3][3];   // B[l][d] = dG_l[d] - dG_3[d]
#pragma unroll
            for (int l = 0; l < 3; ++l)
#pragma unroll
                for (int d = 0; d < 3; ++d) B[l][d] = Qs[144 + 3 * l + d] - Qs[144 + 9 + d];
            // Abar[r][q] = detK * ( e Ai[q][r] - (Ai B^T Ai)[q][r] ),  Ai[q][:] = G[q][:]
            double val = 0.0;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                double t = 0.0;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    double bt = B[0][i] * Gs[r] + B[1][i] * Gs[3 + r] + B[2][i] * Gs[6 + r];   // (B^T Ai)[i][r]
                    t += geo.G[q][i] * bt;
                }
                double aq = geo.detK * (ee * Gs[3 * q + r] - t);
                if (p == q) val += aq;
                if (p == 3) val -= aq;
            }
            val -= s_mass * geo.sgnM * Qs[172 + lane];
            tet_grad[e * 12 + lane] = val;
        }
        __syncwarp();
    }
}

// grad_verts[node] = sum over incident (tet, corner) of tet_grad, cast to fp32
__global__ void k_node_gather(const int32_t* __restrict__ inc_ptr, const int32_t* __restrict__ inc,
                              const double* __restrict__ tet_grad, int64_t n_nodes, float* __restrict__ grad_verts) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= 3 * n_nodes) return;
    int64_t node = t / 3;
    int r = (int)(t - 3 * node);
    double s = 0.0;
    for (int q = inc_ptr[node]; q < inc_ptr[node + 1]; ++q) s += tet_grad[(int64_t)inc[q] * 3 + r];
    grad_verts[t] = (float)s;
}

__global__ void k_corner_keys(const int32_t* __restrict__ tets, int64_t T, int npe, int order,
                              uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= 4 * T) return;
    int64_t e = q >> 2;
    int p = (int)(q & 3);
    int loc = order == 1 ? p : (p == 3 ? 9 : 2 * p);
    keys[q] = (uint32_t)tets[e * npe + loc];
    vals[q] = (uint32_t)q;
}

__global__ void k_inc_ptr(const uint32_t* __restrict__ keys, int64_t nq, int64_t n_nodes, int32_t* __restrict__ inc_ptr) {
    int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= nq) return;
    int64_t cur = keys[q];
    int64_t prev = q == 0 ? -1 : (int64_t)keys[q - 1];
    for (int64_t r = prev + 1; r <= cur; ++r) inc_ptr[r] = (int32_t)q;
    if (q == nq - 1)
        for (int64_t r = cur + 1; r <= n_nodes; ++r) inc_ptr[r] = (int32_t)nq;
}

// ---------------------------------------------------------------------------
// material quadratic forms
// ---------------------------------------------------------------------------
constexpr int QF_WARPS = 4;
constexpr int QF_MAX_CTAS = 148 * 4;

__device__ __forceinline__ void phi_pair(const double F[3][3], double& pmu, double& plam) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) s = fma(F[i][j], F[i][j] + F[j][i], s);
    double tr = F[0][0] + F[1][1] + F[2][2];
    pmu = s;
    plam = tr * tr;
}

template <int ORDER>
__global__ void __launch_bounds__(QF_WARPS * 32)
k_quadforms(const float* __restrict__ verts, const int32_t* __restrict__ tets, int64_t T,
            const double* __restrict__ mtab_g, double wsum, const double* __restrict__ U, int64_t ldu, int k,
            double* __restrict__ partial) {
    constexpr int NPE = ORDER == 1 ? 4 : 10;
    __shared__ double s_mtab[NPE * NPE];
    for (int t = threadIdx.x; t < NPE * NPE; t += blockDim.x) s_mtab[t] = mtab_g[t];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t gw = (int64_t)blockIdx.x * QF_WARPS + warp;
    const int64_t wstride = (int64_t)gridDim.x * QF_WARPS;
    const int kpad = (k + 31) & ~31;
    for (int i0 = 0; i0 < kpad; i0 += 32) {
        const int i = i0 + lane;
        const bool on = i < k;
        double q_mu = 0.0, q_la = 0.0, q_m = 0.0;
        for (int64_t e = gw; e < T; e += wstride) {
            const int32_t* tp = tets + e * NPE;
            int32_t cn[4];
            load_corners(tp, ORDER, cn);
            TetGeom geo;
            tet_geometry<false>(verts, cn, geo);
            double u[NPE][3];
#pragma unroll
            for (int a = 0; a < NPE; ++a) {
                int64_t node = __ldg(tp + a);
#pragma unroll
                for (int c = 0; c < 3; ++c) u[a][c] = on ? __ldg(U + (3 * node + c) * ldu + i) : 0.0;
            }
            // mass form
            double ms = 0.0;
#pragma unroll
            for (int a = 0; a < NPE; ++a) {
                double ta[3] = {0.0, 0.0, 0.0};
#pragma unroll
                for (int b = a + 1; b < NPE; ++b) {
                    double mab = s_mtab[a * NPE + b];
                    ta[0] = fma(mab, u[b][0], ta[0]);
                    ta[1] = fma(mab, u[b][1], ta[1]);
                    ta[2] = fma(mab, u[b][2], ta[2]);
                }
                double maa = s_mtab[a * NPE + a];
#pragma unroll
                for (int c = 0; c < 3; ++c) ms = fma(u[a][c], fma(maa, u[a][c], 2.0 * ta[c]), ms);
            }
            q_m = fma(geo.detM, ms, q_m);
            if (ORDER == 1) {
                double F[3][3];
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        F[r][j] = u[0][r] * geo.G[0][j] + u[1][r] * geo.G[1][j] + u[2][r] * geo.G[2][j] +
                                  u[3][r] * geo.G[3][j];
                double pm, pl;
                phi_pair(F, pm, pl);
                q_mu = fma(geo.detK * wsum, pm, q_mu);
                q_la = fma(geo.detK * wsum, pl, q_la);
            } else {
                // corner / edge nodes in the reference local order
                constexpr int CORNER[4] = {0, 2, 4, 9};
                constexpr int EDGE[4][4] = {{-1, 1, 5, 6}, {1, -1, 3, 7}, {5, 3, -1, 8}, {6, 7, 8, -1}};
                double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
                double smu = 0.0, sla = 0.0;
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    double F[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
                    for (int l = 0; l < 4; ++l) {
                        double h[3];
#pragma unroll
                        for (int r = 0; r < 3; ++r)
                            h[r] = (l == m) ? 3.0 * u[CORNER[l]][r] : 4.0 * u[EDGE[l][m] < 0 ? 0 : EDGE[l][m]][r] - u[CORNER[l]][r];
#pragma unroll
                        for (int r = 0; r < 3; ++r)
#pragma unroll
                            for (int j = 0; j < 3; ++j) F[r][j] = fma(h[r], geo.G[l][j], F[r][j]);
                    }
                    double pm, pl;
                    phi_pair(F, pm, pl);
                    smu += pm;
                    sla += pl;
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int j = 0; j < 3; ++j) S[r][j] += F[r][j];
                }
                double pm, pl;
                phi_pair(S, pm, pl);
                const double w = geo.detK * (1.0 / 120.0);
                q_mu = fma(w, smu + pm, q_mu);
                q_la = fma(w, sla + pl, q_la);
            }
        }
        if (on) {
            double* out = partial + gw * 3 * (int64_t)k;
            out[i] = q_mu;
            out[k + i] = q_la;
            out[2 * k + i] = q_m;
        }
    }
}

__global__ void k_quadform_reduce(const double* __restrict__ partial, int nparts, int width, double* __restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= width) return;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += partial[(size_t)p * width + t];
    out[t] = s;
}

static int qf_ctas(int64_t T) {
    int64_t c = ceil_div(T, QF_WARPS * 8);
    if (c < 1) c = 1;
    return (int)(c < QF_MAX_CTAS ? c : QF_MAX_CTAS);
}

}  // namespace ds

using namespace ds;

extern "C" int ds_corner_incidence(ds_workspace* ws, const int32_t* tets, int64_t T, int npe, int order,
                                   int64_t n_nodes, int32_t* inc_ptr, int32_t* inc, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DS_REQUIRE(ws && tets && inc_ptr && inc, "ds_corner_incidence: null argument");
    DS_REQUIRE((order == 1 && npe == 4) || (order == 2 && npe == 10), "ds_corner_incidence: order/npe mismatch");
    DS_REQUIRE(T > 0 && n_nodes > 0 && 4 * T < (int64_t)2147483647, "ds_corner_incidence: bad sizes");
    int64_t nq = 4 * T;
    ProfScope prof(PROF_PATTERN, stream);
    int end_bit = 1;
    while (end_bit < 32 && ((int64_t)1 << end_bit) < n_nodes) ++end_bit;
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)nq, 0, end_bit, stream);
    DS_TRY(ws->arena.reserve((size_t)nq * 12 + tmp + 8 * 256, stream));
    uint32_t* keys_in = ws->arena.take<uint32_t>(nq);
    uint32_t* keys_out = ws->arena.take<uint32_t>(nq);
    uint32_t* vals_in = ws->arena.take<uint32_t>(nq);
    void* cub_tmp = ws->arena.take<char>(tmp);
    DS_REQUIRE(cub_tmp != nullptr, "ds_corner_incidence: arena too small");
    unsigned blocks = (unsigned)ceil_div(nq, 256);
    k_corner_keys<<<blocks, 256, 0, stream>>>(tets, T, npe, order, keys_in, vals_in);
    DS_LAUNCH_CHECK();
    DS_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, tmp, keys_in, keys_out, vals_in, (uint32_t*)inc, (int)nq, 0,
                                            end_bit, stream));
    k_inc_ptr<<<blocks, 256, 0, stream>>>(keys_out, nq, n_nodes, inc_ptr);
    DS_LAUNCH_CHECK();
    return DS_OK;
}

extern "C" int ds_eigval_grad_shape(const float* verts, const int32_t* tets, int64_t T, int order, int64_t n_nodes,
                                    double mu, double lam_lame, const double* ctab, const double* mtab,
                                    const double* U, int64_t ldu, int k, const double* lam, const double* g,
                                    const int32_t* inc_ptr, const int32_t* inc, double* tet_grad,
                                    float* grad_verts, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DS_REQUIRE(order == 1 || order == 2, "ds_eigval_grad_shape: order must be 1 or 2 (got %d)", order);
    DS_REQUIRE(verts && tets && ctab && mtab && U && lam && g && inc_ptr && inc && tet_grad && grad_verts,
               "ds_eigval_grad_shape: null argument");
    DS_REQUIRE(T > 0 && n_nodes > 0 && k > 0 && k <= 1024 && ldu >= k, "ds_eigval_grad_shape: bad sizes");
    const int kpad = (k + 15) & ~15;
    ProfScope prof(PROF_GRAD, stream);
    {   // n k 8 (U once) + T npe 4 + nodes 12 + nodes 24; flops: per tet and mode the quadratic forms of the 4 corner
        // derivatives, ~ 2 (3 npe)^2 x 12 / 4 (counted from the kernel's DMMA tiles in DESIGN.md)
        const int npe_ = order == 1 ? 4 : 10;
        prof_account(PROF_GRAD, (double)n_nodes * 3.0 * k * 8.0 + (double)T * npe_ * 4.0 + (double)n_nodes * 36.0, 0.0);
    }
    int64_t ctas64 = ceil_div(T, GS_WARPS);
    int ctas = (int)(ctas64 < 148 * 8 ? ctas64 : 148 * 8);
    if (order == 1) {
        size_t smem = sizeof(GradSmem<1>) + 2 * kpad * sizeof(double);
        DS_CUDA(cudaFuncSetAttribute(k_eigval_grad_shape<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_eigval_grad_shape<1><<<ctas, GS_WARPS * 32, smem, stream>>>(verts, tets, T, mu, lam_lame, ctab, mtab, U, ldu, k,
                                                                      lam, g, tet_grad);
    } else {
        size_t smem = sizeof(GradSmem<2>) + 2 * kpad * sizeof(double);
        DS_CUDA(cudaFuncSetAttribute(k_eigval_grad_shape<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_eigval_grad_shape<2><<<ctas, GS_WARPS * 32, smem, stream>>>(verts, tets, T, mu, lam_lame, ctab, mtab, U, ldu, k,
                                                                      lam, g, tet_grad);
    }
    DS_LAUNCH_CHECK();
    k_node_gather<<<(unsigned)ceil_div(3 * n_nodes, 256), 256, 0, stream>>>(inc_ptr, inc, tet_grad, n_nodes, grad_verts);
    DS_LAUNCH_CHECK();
    return DS_OK;
}

extern "C" int64_t ds_quadform_scratch_elems(int k) { return (int64_t)QF_MAX_CTAS * QF_WARPS * 3 * k; }

extern "C" int ds_eigval_quadforms_material(const float* verts, const int32_t* tets, int64_t T, int order,
                                            const double* mtab, double wsum, const double* U, int64_t ldu, int k,
                                            double* partial, double* out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DS_REQUIRE(order == 1 || order == 2, "ds_eigval_quadforms_material: order must be 1 or 2 (got %d)", order);
    DS_REQUIRE(verts && tets && mtab && U && partial && out, "ds_eigval_quadforms_material: null argument");
    DS_REQUIRE(T > 0 && k > 0 && ldu >= k, "ds_eigval_quadforms_material: bad sizes");
    int ctas = qf_ctas(T);
    ProfScope prof(PROF_QUADFORM, stream);
    if (order == 1)
        k_quadforms<1><<<ctas, QF_WARPS * 32, 0, stream>>>(verts, tets, T, mtab, wsum, U, ldu, k, partial);
    else
        k_quadforms<2><<<ctas, QF_WARPS * 32, 0, stream>>>(verts, tets, T, mtab, wsum, U, ldu, k, partial);
    DS_LAUNCH_CHECK();
    k_quadform_reduce<<<(3 * k + 127) / 128, 128, 0, stream>>>(partial, ctas * QF_WARPS, 3 * k, out);
    DS_LAUNCH_CHECK();
    return DS_OK;
}
