// Fused stiffness + mass assembly into the fixed block-CSR pattern.
//
// Reference behaviour replaced (file:line under /root/reference/src):
//   diffelastic/diff_model.py:184-220  update_stiff_matrix  (batched A^T B A + COO coalesce)
//   diffelastic/diff_model.py:222-312  update_mass_matrix   (900 strided writes + coalesce)
//   diffelastic/deform.py:35-68,136-147  grad_x N and w_g |det A| tables (T*G*npe*3 fp32)
//   cuda/massMatrixDouble.cu:3-78       compute_mass_matrix_kernel (1 thread / tet, COO)
//
// Design: geometry is affine per tet (mesh.py:69-99 uses the corner nodes only), so
//   grad_x N_a(g) = sum_l dN_a/dL_l(g) G_l,   G = dL/dxi A^-1   (4x3 per tet)
// and the Gauss sum collapses into a constant table
//   ctab[a][b][l][m] = sum_g w_g dN_a/dL_l(g) dN_b/dL_m(g)
// built on the host from the reference's own fp32 rule.  Then per element block
//   S_ab = sum_{l,m} ctab[a][b][l][m] G_l G_m^T
//   K_ab[c][d] = |det A| ( mu (delta_cd tr S_ab + S_ab[d][c]) + lam S_ab[c][d] )
//   M_ab = mtab[a][b] |det6V| I3.
// Owner-computes: lane p of the warp that owns node row i sums all element
// contributions of block (i, bcol[brow[i]+p]) in ascending element order.
#include "common.cuh"
#include "../../include/diffsound_sm100.h"

namespace ds {

constexpr int GEOM_STRIDE = 14;  // G[4][3], detK, detM

__device__ __constant__ int c_sup2[10][2] = {{0, 0}, {0, 1}, {1, 1}, {1, 2}, {2, 2},
                                             {2, 0}, {0, 3}, {1, 3}, {2, 3}, {3, 3}};
__device__ __constant__ int c_nsup2[10] = {1, 2, 1, 2, 1, 2, 2, 2, 2, 1};

__device__ __forceinline__ void corner_ids(const int32_t* t, int order, int32_t c[4]) {
    if (order == 1) {
        c[0] = t[0]; c[1] = t[1]; c[2] = t[2]; c[3] = t[3];
    } else {
        c[0] = t[0]; c[1] = t[2]; c[2] = t[4]; c[3] = t[9];
    }
}

// One thread per tet.  A is built from fp32 differences exactly as mesh.py:90-98;
// inverse and determinant are then taken in fp64 (more accurate than the
// reference's fp32 torch.inverse / torch.det).
__global__ void k_tet_geometry(const float* __restrict__ verts, const int32_t* __restrict__ tets, int64_t T,
                               int npe, int order, double* __restrict__ geom) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= T) return;
    int32_t c[4];
    corner_ids(tets + e * npe, order, c);
    float v[4][3];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        v[p][0] = verts[3 * (int64_t)c[p] + 0];
        v[p][1] = verts[3 * (int64_t)c[p] + 1];
        v[p][2] = verts[3 * (int64_t)c[p] + 2];
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
    double Ai[3][3];
    Ai[0][0] = c00 * id;
    Ai[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) * id;
    Ai[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) * id;
    Ai[1][0] = c01 * id;
    Ai[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * id;
    Ai[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) * id;
    Ai[2][0] = c02 * id;
    Ai[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) * id;
    Ai[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * id;
    double* g = geom + e * GEOM_STRIDE;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        g[0 + d] = Ai[0][d];
        g[3 + d] = Ai[1][d];
        g[6 + d] = Ai[2][d];
        g[9 + d] = -(Ai[0][d] + Ai[1][d] + Ai[2][d]);
    }
    g[12] = fabs(det);
    // |6V| from fp64 corner coordinates, same expansion as diff_model.py:272-289
    // (no FMA contraction so that it rounds like the CPU reference).
    double x[4], y[4], z[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        x[p] = (double)v[p][0];
        y[p] = (double)v[p][1];
        z[p] = (double)v[p][2];
    }
    double x10 = __dsub_rn(x[1], x[0]), y10 = __dsub_rn(y[1], y[0]), z10 = __dsub_rn(z[1], z[0]);
    double x20 = __dsub_rn(x[2], x[0]), y20 = __dsub_rn(y[2], y[0]), z20 = __dsub_rn(z[2], z[0]);
    double x30 = __dsub_rn(x[3], x[0]), y30 = __dsub_rn(y[3], y[0]), z30 = __dsub_rn(z[3], z[0]);
    double t1 = __dmul_rn(x10, __dsub_rn(__dmul_rn(y20, z30), __dmul_rn(y30, z20)));
    double t2 = __dmul_rn(y10, __dsub_rn(__dmul_rn(z20, x30), __dmul_rn(z30, x20)));
    double t3 = __dmul_rn(z10, __dsub_rn(__dmul_rn(x20, y30), __dmul_rn(x30, y20)));
    g[13] = fabs(__dadd_rn(__dadd_rn(t1, t2), t3));
}

template <int ORDER>
__global__ void __launch_bounds__(256)
k_assemble_rows(const double* __restrict__ geom, const double* __restrict__ ctab_g,
                const double* __restrict__ mtab_g, const int32_t* __restrict__ brow,
                const int32_t* __restrict__ contrib_ptr, const int32_t* __restrict__ contrib,
                int64_t n_nodes, double mu, double lam, double* __restrict__ Kval, double* __restrict__ Mblk) {
    constexpr int NPE = ORDER == 1 ? 4 : 10;
    constexpr int NPE2 = NPE * NPE;
    __shared__ double s_ctab[NPE2 * 16];
    __shared__ double s_mtab[NPE2];
    for (int t = threadIdx.x; t < NPE2 * 16; t += blockDim.x) s_ctab[t] = ctab_g[t];
    for (int t = threadIdx.x; t < NPE2; t += blockDim.x) s_mtab[t] = mtab_g[t];
    __syncthreads();
    int lane = threadIdx.x & 31;
    int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= n_nodes) return;
    int64_t b0 = brow[row];
    int deg = (int)(brow[row + 1] - b0);
    for (int p = lane; p < deg; p += 32) {
        int64_t s = b0 + p;
        double acc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        double macc = 0.0;
        int q0 = contrib_ptr[s], q1 = contrib_ptr[s + 1];
        for (int q = q0; q < q1; ++q) {
            int pair = contrib[q];
            int e = pair / NPE2;
            int ab = pair - e * NPE2;
            int a = ab / NPE, b = ab - a * NPE;
            const double* g = geom + (int64_t)e * GEOM_STRIDE;
            double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
            int na, nb, la[2], lb[2];
            if (ORDER == 1) {
                na = nb = 1;
                la[0] = a; lb[0] = b; la[1] = lb[1] = 0;
            } else {
                na = c_nsup2[a]; nb = c_nsup2[b];
                la[0] = c_sup2[a][0]; la[1] = c_sup2[a][1];
                lb[0] = c_sup2[b][0]; lb[1] = c_sup2[b][1];
            }
            const double* ct = s_ctab + ab * 16;
            for (int li = 0; li < na; ++li) {
                int l = la[li];
                double gl0 = g[3 * l], gl1 = g[3 * l + 1], gl2 = g[3 * l + 2];
                for (int mi = 0; mi < nb; ++mi) {
                    int m = lb[mi];
                    double c = ct[l * 4 + m];
                    double gm0 = g[3 * m], gm1 = g[3 * m + 1], gm2 = g[3 * m + 2];
                    double a0 = c * gl0, a1 = c * gl1, a2 = c * gl2;
                    S[0][0] += a0 * gm0; S[0][1] += a0 * gm1; S[0][2] += a0 * gm2;
                    S[1][0] += a1 * gm0; S[1][1] += a1 * gm1; S[1][2] += a1 * gm2;
                    S[2][0] += a2 * gm0; S[2][1] += a2 * gm1; S[2][2] += a2 * gm2;
                }
            }
            double detK = g[12], detM = g[13];
            double tr = mu * (S[0][0] + S[1][1] + S[2][2]);
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    double v = mu * S[d][c] + lam * S[c][d];
                    if (c == d) v += tr;
                    acc[c][d] += detK * v;
                }
            macc += s_mtab[ab] * detM;
        }
        double* out = Kval + 9 * b0 + 3 * p;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int d = 0; d < 3; ++d) out[(int64_t)c * 3 * deg + d] = acc[c][d];
        Mblk[s] = macc;
    }
}

__global__ void k_mass_expand(const int32_t* __restrict__ brow, int64_t n_nodes, const double* __restrict__ Mblk,
                              double* __restrict__ Mval) {
    int lane = threadIdx.x & 31;
    int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= n_nodes) return;
    int64_t b0 = brow[row];
    int64_t deg = brow[row + 1] - b0;
    for (int64_t t = lane; t < 9 * deg; t += 32) {
        int c = (int)(t / (3 * deg));
        int64_t r = t - c * 3 * deg;
        int64_t p = r / 3;
        int d = (int)(r - 3 * p);
        Mval[9 * b0 + t] = (c == d) ? Mblk[b0 + p] : 0.0;
    }
}

// Legacy COO export, one thread per output triple (coalesced writes; the
// reference kernel writes msize^2 strided entries per thread).
__global__ void k_mass_coo(const double* __restrict__ vertices, const int32_t* __restrict__ tets, int64_t T,
                           int order, int vnum, const double* __restrict__ m, double d,
                           double* __restrict__ values, int32_t* __restrict__ rows, int32_t* __restrict__ cols) {
    int msize = 3 * vnum;
    int64_t msz2 = (int64_t)msize * msize;
    int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= T * msz2) return;
    int64_t e = idx / msz2;
    int r = (int)(idx - e * msz2);
    int i = r / msize, j = r - i * msize;
    const int32_t* t = tets + e * vnum;
    int32_t c[4];
    if (order == 1) { c[0] = t[0]; c[1] = t[1]; c[2] = t[2]; c[3] = t[3]; }
    else if (order == 2) { c[0] = t[0]; c[1] = t[2]; c[2] = t[4]; c[3] = t[9]; }
    else { c[0] = t[0]; c[1] = t[3]; c[2] = t[6]; c[3] = t[16]; }
    double x[4], y[4], z[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        x[p] = vertices[3 * (int64_t)c[p]];
        y[p] = vertices[3 * (int64_t)c[p] + 1];
        z[p] = vertices[3 * (int64_t)c[p] + 2];
    }
    double V = ((x[1] - x[0]) * ((y[2] - y[0]) * (z[3] - z[0]) - (y[3] - y[0]) * (z[2] - z[0])) +
                (y[1] - y[0]) * ((x[3] - x[0]) * (z[2] - z[0]) - (x[2] - x[0]) * (z[3] - z[0])) +
                (z[1] - z[0]) * ((x[2] - x[0]) * (y[3] - y[0]) - (x[3] - x[0]) * (y[2] - y[0]))) / 6;
    V = fabs(V) * 6;
    values[idx] = m[r] * d * V;
    rows[idx] = t[i / 3] * 3 + i % 3;
    cols[idx] = t[j / 3] * 3 + j % 3;
}

}  // namespace ds

using namespace ds;

extern "C" int ds_assemble_km(const float* verts, const int32_t* tets, int64_t T, int order, int64_t n_nodes,
                              double mu, double lam, const double* ctab, const double* mtab,
                              const int32_t* brow, const int32_t* bcol, const int32_t* contrib_ptr,
                              const int32_t* contrib, int64_t nnzb, double* geom, double* Kval, double* Mblk,
                              void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    (void)bcol; (void)nnzb;
    DS_REQUIRE(order == 1 || order == 2, "ds_assemble_km: order must be 1 or 2 (got %d)", order);
    DS_REQUIRE(verts && tets && ctab && mtab && brow && contrib_ptr && contrib && geom && Kval && Mblk,
               "ds_assemble_km: null argument");
    DS_REQUIRE(T > 0 && n_nodes > 0, "ds_assemble_km: empty mesh");
    int npe = order == 1 ? 4 : 10;
    ProfScope prof(PROF_ASSEMBLE, stream);
    k_tet_geometry<<<(unsigned)ceil_div(T, 128), 128, 0, stream>>>(verts, tets, T, npe, order, geom);
    DS_LAUNCH_CHECK();
    unsigned blocks = (unsigned)ceil_div(n_nodes * 32, 256);
    if (order == 1)
        k_assemble_rows<1><<<blocks, 256, 0, stream>>>(geom, ctab, mtab, brow, contrib_ptr, contrib, n_nodes, mu,
                                                       lam, Kval, Mblk);
    else
        k_assemble_rows<2><<<blocks, 256, 0, stream>>>(geom, ctab, mtab, brow, contrib_ptr, contrib, n_nodes, mu,
                                                       lam, Kval, Mblk);
    DS_LAUNCH_CHECK();
    return DS_OK;
}

extern "C" int ds_mass_expand(const int32_t* brow, int64_t n_nodes, int64_t nnzb, const double* Mblk,
                              double* Mval, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    (void)nnzb;
    DS_REQUIRE(brow && Mblk && Mval, "ds_mass_expand: null argument");
    k_mass_expand<<<(unsigned)ceil_div(n_nodes * 32, 256), 256, 0, stream>>>(brow, n_nodes, Mblk, Mval);
    DS_LAUNCH_CHECK();
    return DS_OK;
}

extern "C" int ds_assemble_mass_coo(const double* vertices, const int32_t* tets, int64_t T, int order,
                                    const double* element_mm, double density, double* values, int32_t* rows,
                                    int32_t* cols, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DS_REQUIRE(order >= 1 && order <= 3, "ds_assemble_mass_coo: order must be 1, 2 or 3 (got %d)", order);
    DS_REQUIRE(vertices && tets && element_mm && values && rows && cols, "ds_assemble_mass_coo: null argument");
    if (T == 0) return DS_OK;
    int vnum = order == 1 ? 4 : (order == 2 ? 10 : 20);
    int64_t total = T * 9 * vnum * vnum;
    k_mass_coo<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(vertices, tets, T, order, vnum, element_mm,
                                                                  density, values, rows, cols);
    DS_LAUNCH_CHECK();
    return DS_OK;
}
