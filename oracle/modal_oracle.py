"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the modal-analysis hot path.

A CPU restatement (torch-CPU / numpy / scipy) of the reference's algorithm for
the path  mesh -> K, M -> lowest eigenpairs -> d(lambda)/d(theta) -> audio.
Every function cites the reference file:line it follows.  Nothing under
``diffsound_b200/`` may import this module; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs use it, as the checker / reported CPU baseline.

Pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so this oracle is pinned against outputs of the
reference itself, run unmodified in the build container through
``oracle/ref_harness.py`` and committed as ``tests/golden/*.npz`` by
``oracle/make_goldens.py`` (see tests/test_oracle_vs_golden.py).

Precision ladder kept as in the reference: fp32 vertices,
transform matrix, inverse, shape-function gradients and integration weights;
fp64 contraction, matrix values, eigen-solve.
"""
import itertools
import math

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch

NPE = {1: 4, 2: 10}
# local ids of the four corner nodes inside an element (mesh.py:75-89)
CORNERS = {1: (0, 1, 2, 3), 2: (0, 2, 4, 9)}


# ----------------------------------------------------------------------------
# quadrature and shape functions
# ----------------------------------------------------------------------------
def gauss_rule(n):
    """Collapsed-cube Gauss-Legendre rule on the unit tet, n^3 points, fp32.
    Follows src/diffelastic/gauss.py:4-37 (roots of P_n via numpy legroots,
    weights 2/((1-r^2) P_n'(r)^2), points mapped to [0,1], Duffy collapse)."""
    from numpy.polynomial.legendre import Legendre, legroots
    c = np.zeros(n + 1, dtype=np.float32)
    c[-1] = 1
    poly = Legendre(c)
    r = legroots(c)
    dv = poly.deriv()(r)
    wt = 2 / ((1 - r ** 2) * dv ** 2)
    r = (r + 1) / 2
    pts = np.zeros((n ** 3, 4), dtype=np.float32)
    wts = np.zeros(n ** 3, dtype=np.float32)
    for i, j, k in itertools.product(range(n), repeat=3):
        q = i * n * n + j * n + k
        w = np.float32(r[i])
        z = np.float32(r[j] * (1 - w))
        y = np.float32(r[k] * (1 - w - z))
        x = np.float32(1 - w - z - y)
        pts[q] = (x, y, z, w)
        wts[q] = wt[i] * wt[j] * wt[k] * (1 - w) * (1 - w - z) / 8
    return pts, wts


def shape_fn(L, order):
    """N_a(L) for 4- / 10-node tets (src/diffelastic/shape_func.py:3-24).
    L: (G,4) torch tensor; returns (G, npe)."""
    L1, L2, L3, L4 = L[:, 0], L[:, 1], L[:, 2], L[:, 3]
    if order == 1:
        return L
    cols = [L1 * (2 * L1 - 1), 4 * L1 * L2, L2 * (2 * L2 - 1), 4 * L2 * L3,
            L3 * (2 * L3 - 1), 4 * L3 * L1, 4 * L1 * L4, 4 * L2 * L4,
            4 * L3 * L4, L4 * (2 * L4 - 1)]
    return torch.stack(cols, dim=1)


def shape_fn_grad(L, order):
    """dN_a/dL_l (src/diffelastic/shape_func.py:51-84).  Returns (G, npe, 4)."""
    G = L.shape[0]
    out = torch.zeros(G, NPE[order], 4, dtype=L.dtype)
    if order == 1:
        for a in range(4):
            out[:, a, a] = 1
        return out
    L1, L2, L3, L4 = L[:, 0], L[:, 1], L[:, 2], L[:, 3]
    one = torch.ones_like(L1)
    out[:, 0, 0] = 4 * L1 - one
    out[:, 1, 0], out[:, 1, 1] = 4 * L2, 4 * L1
    out[:, 2, 1] = 4 * L2 - one
    out[:, 3, 1], out[:, 3, 2] = 4 * L3, 4 * L2
    out[:, 4, 2] = 4 * L3 - one
    out[:, 5, 0], out[:, 5, 2] = 4 * L3, 4 * L1
    out[:, 6, 0], out[:, 6, 3] = 4 * L4, 4 * L1
    out[:, 7, 1], out[:, 7, 3] = 4 * L4, 4 * L2
    out[:, 8, 2], out[:, 8, 3] = 4 * L4, 4 * L3
    out[:, 9, 3] = 4 * L4 - one
    return out


def element_mass_table(order):
    """Reference-element consistent mass  int N_a N_b  by the (order+2)^3 rule,
    fp32, shape (npe, npe) (src/diffelastic/mass_matrix.py:9-23)."""
    pts, wts = gauss_rule(order + 2)
    pts = torch.from_numpy(pts)
    wts = torch.from_numpy(wts)
    N = shape_fn(pts, order)
    npe = NPE[order]
    M = torch.zeros(npe, npe, dtype=torch.float32)
    for a in range(npe):
        for b in range(npe):
            M[a, b] = torch.sum(N[:, a] * N[:, b] * wts)
    return M


# ----------------------------------------------------------------------------
# mesh
# ----------------------------------------------------------------------------
def promote(verts, tets, order):
    """Linear -> quadratic promotion with lexicographic de-duplication.
    verts (V,3) fp32 torch, tets (T,4) int64 torch.  Follows
    src/diffelastic/mesh.py:101-179: six mid-edge points per tet in the order
    (01,12,02,03,13,23), local node order [v0,m01,v1,m12,v2,m02,m03,m13,m23,v3],
    then torch.unique(dim=0) renumbering; representative of each new node =
    smallest original index (scatter-min, mesh.py:176).  Order 1 is returned
    untouched (mesh.py:111-112)."""
    if order == 1:
        return verts, tets
    T = tets.shape[0]
    V = verts.shape[0]
    vf = verts[tets]
    a, b, c, d = vf[:, 0], vf[:, 1], vf[:, 2], vf[:, 3]
    mids = torch.cat([(a + b) / 2, (b + c) / 2, (a + c) / 2,
                      (a + d) / 2, (b + d) / 2, (c + d) / 2], dim=0)
    allv = torch.cat([verts, mids], dim=0)
    ar = torch.arange(T, dtype=tets.dtype)
    nt = torch.stack([tets[:, 0], V + ar, tets[:, 1], V + T + ar, tets[:, 2],
                      V + 2 * T + ar, V + 3 * T + ar, V + 4 * T + ar,
                      V + 5 * T + ar, tets[:, 3]], dim=1)
    _, inv = torch.unique(allv.detach(), dim=0, return_inverse=True)
    nnew = int(inv.max()) + 1
    first = torch.full((nnew,), allv.shape[0], dtype=torch.long)
    first.scatter_reduce_(0, inv, torch.arange(allv.shape[0]), "amin")
    return allv[first], inv[nt]


def transform_matrix(verts, tets, order):
    """A = [x1-x4, x2-x4, x3-x4] columns per tet, fp32 (mesh.py:69-99)."""
    c = CORNERS[order]
    v1, v2, v3, v4 = (verts[tets[:, i]] for i in c)
    return torch.stack([v1 - v4, v2 - v4, v3 - v4], dim=2).to(torch.float32)


# The reference inverts A and takes det A in fp32 (torch.inverse / torch.det on the fp32 transform matrix).  INVERSE_DTYPE =
# torch.float64 evaluates the SAME formulas with the inverse and the determinant in fp64 on the same fp32 A -- not the
# reference's arithmetic; the tests use it only to measure how far the reference's fp32 inverse is from the exact one on a
# given mesh (its own accuracy floor: thin shell elements are ill-conditioned), see `inverse_precision`.
INVERSE_DTYPE = torch.float32


class inverse_precision:
    """with inverse_precision(torch.float64): ... -- A^-1 and det A of the element geometry in that precision."""

    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        global INVERSE_DTYPE
        self.prev, INVERSE_DTYPE = INVERSE_DTYPE, self.dtype

    def __exit__(self, *exc):
        global INVERSE_DTYPE
        INVERSE_DTYPE = self.prev


def shape_func_deriv(verts, tets, order):
    """grad_x N_a at every Gauss point: (T, G, npe, 3) fp32
    = (dN/dL . dL/dxi) . A^-1   (src/diffelastic/deform.py:35-68)."""
    A = transform_matrix(verts, tets, order)
    Ainv = torch.inverse(A.to(INVERSE_DTYPE))
    pts, _ = gauss_rule(order + 2)
    dN = shape_fn_grad(torch.from_numpy(pts), order)          # (G,npe,4) fp32
    dLdxi = torch.tensor([[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, -1, -1]],
                         dtype=torch.float32)
    dNdxi = (dN @ dLdxi).to(INVERSE_DTYPE)                      # (G,npe,3)
    return dNdxi.unsqueeze(0) @ Ainv.unsqueeze(1)               # (T,G,npe,3)


def integration_weights(verts, tets, order):
    """w_g |det A|, (T, G) fp32 (src/diffelastic/deform.py:136-147)."""
    A = transform_matrix(verts, tets, order)
    _, wts = gauss_rule(order + 2)
    return torch.abs(torch.det(A.to(INVERSE_DTYPE))).unsqueeze(1) * torch.from_numpy(wts).to(INVERSE_DTYPE).unsqueeze(0)


def lame(youngs, poisson):
    """(mu, lambda) as in FixedLinear.get_stress (diff_model.py:34-41)."""
    lam = youngs * poisson / ((1 + poisson) * (1 - 2 * poisson))
    mu = youngs / (2 * (1 + poisson))
    return mu, lam


# ----------------------------------------------------------------------------
# element matrices (differentiable in verts through torch autograd)
# ----------------------------------------------------------------------------
def element_stiffness(verts, tets, order, mu, lam, sl=None):
    """(Tb, 3npe, 3npe) fp64 element stiffness = sum_g w (A^T B A) with
    B = d(stress)/dF of P = mu(F+F^T) + lam tr(F) I  (diff_model.py:39-41,
    184-213): K[3a+c,3b+d] = w(mu(delta_cd gNa.gNb + gNa[d] gNb[c]) + lam gNa[c] gNb[d])."""
    t = tets if sl is None else tets[sl]
    g = shape_func_deriv(verts, t, order).double()              # (T,G,N,3)
    w = integration_weights(verts, t, order).double()           # (T,G)
    T, G, N, _ = g.shape
    gw = g * w[:, :, None, None]
    dot = torch.einsum("tgac,tgbc->tab", gw, g)                 # sum_g w gNa.gNb
    outer = torch.einsum("tgac,tgbd->tacbd", gw, g)             # sum_g w gNa[c] gNb[d]
    eye = torch.eye(3, dtype=torch.float64)
    K = (mu * (dot[:, :, None, :, None] * eye[None, None, :, None, :]
               + outer.permute(0, 1, 4, 3, 2))
         + lam * outer)
    return K.reshape(T, 3 * N, 3 * N)


def tet_det64(verts, tets, order):
    """|6V| from fp64 corner coordinates with the reference's expansion
    (diff_model.py:233-289; same expression as src/cuda/massMatrixDouble.cu:59-60)."""
    c = CORNERS[order]
    p = [verts[tets[:, i]].double() for i in c]
    x = [q[:, 0] for q in p]
    y = [q[:, 1] for q in p]
    z = [q[:, 2] for q in p]
    V = ((x[1] - x[0]) * ((y[2] - y[0]) * (z[3] - z[0]) - (y[3] - y[0]) * (z[2] - z[0]))
         + (y[1] - y[0]) * ((z[2] - z[0]) * (x[3] - x[0]) - (z[3] - z[0]) * (x[2] - x[0]))
         + (z[1] - z[0]) * ((x[2] - x[0]) * (y[3] - y[0]) - (x[3] - x[0]) * (y[2] - y[0])))
    return torch.abs(V)


def element_mass(verts, tets, order, density, sl=None):
    """(Tb, 3npe, 3npe) fp64: (fp32 table * density rounded to fp32) * |6V|
    (diff_model.py:299-303, mass_matrix.py:25-31)."""
    t = tets if sl is None else tets[sl]
    tab = (element_mass_table(order) * density).double()        # fp32 product, then promoted
    V = tet_det64(verts, t, order)
    npe = NPE[order]
    Me = tab[None] * V[:, None, None]
    eye = torch.eye(3, dtype=torch.float64)
    return (Me[:, :, None, :, None] * eye[None, None, :, None, :]).reshape(-1, 3 * npe, 3 * npe)


def element_dofs(tets):
    """dof id 3*node+c per (tet, node, c) (deform.py:113-125)."""
    return (tets[:, :, None] * 3 + torch.arange(3)[None, None, :]).reshape(tets.shape[0], -1)


def assemble(verts, tets, order, youngs, poisson, density, batch=4096):
    """K, M as scipy CSR on the reference's coalesced-COO pattern, entries in
    (row, col) order, explicit zeros kept (diff_model.py:216-220, 311-312;
    SURVEY A.3).  The reference sums duplicates by sort-and-reduce per batch;
    here every element entry is binned straight into its pattern slot."""
    n = 3 * verts.shape[0]
    mu, lam = lame(youngs, poisson)
    crow, col, _, _ = pattern(tets, verts.shape[0])
    rows_of = np.repeat(np.arange(n, dtype=np.int64), np.diff(crow))
    keys = rows_of * n + col
    kv = np.zeros(col.size)
    mv = np.zeros(col.size)
    for s in range(0, tets.shape[0], batch):
        sl = slice(s, min(s + batch, tets.shape[0]))
        d = element_dofs(tets[sl]).numpy().astype(np.int64)
        T, m = d.shape
        ek = (d[:, :, None] * n + d[:, None, :]).reshape(-1)
        slot = np.searchsorted(keys, ek)
        Ke = element_stiffness(verts, tets, order, mu, lam, sl).detach().numpy().reshape(-1)
        Me = element_mass(verts, tets, order, density, sl).detach().numpy().reshape(-1)
        kv += np.bincount(slot, weights=Ke, minlength=col.size)
        mv += np.bincount(slot, weights=Me, minlength=col.size)
    K = sp.csr_matrix((kv, col, crow), shape=(n, n))
    M = sp.csr_matrix((mv, col, crow), shape=(n, n))
    return K, M


def pattern(tets, n_nodes):
    """Bit-exact CSR pattern (crow int64, col int64) of the coalesced K/M
    (diff_model.py:216-220): dense 3x3 blocks over ordered node
    pairs sharing an element, sorted by (row, col).  Also returns the
    node-level block CSR (brow, bcol)."""
    t = tets.numpy().astype(np.int64)
    keys = np.unique((t[:, :, None] * n_nodes + t[:, None, :]).reshape(-1))
    bi, bj = keys // n_nodes, keys % n_nodes
    deg = np.bincount(bi, minlength=n_nodes)
    brow = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    crow = np.concatenate([[0], np.cumsum(np.repeat(3 * deg, 3))]).astype(np.int64)
    p = np.arange(keys.size) - brow[bi]
    col = np.empty(9 * keys.size, dtype=np.int64)
    for c in range(3):
        for d in range(3):
            col[9 * brow[bi] + c * 3 * deg[bi] + 3 * p + d] = 3 * bj + d
    return crow, col, brow, bj.astype(np.int64)


# ----------------------------------------------------------------------------
# eigen-solve and eigenvalue derivative
# ----------------------------------------------------------------------------
def eig_arpack(K, M, k, sigma=20000.0):
    """k+6 eigenpairs nearest sigma by shift-invert ARPACK, rigid six dropped
    (diff_model.py:335-369).  Returns (lam (k,), U (n,k), U_full (n,k+6), S)."""
    Kz = K.copy()
    Kz.eliminate_zeros()
    Mz = M.copy()
    Mz.eliminate_zeros()
    S, U = spla.eigsh(Kz, M=Mz, k=k + 6, sigma=sigma)
    return S[6:], U[:, 6:6 + k], U, S


def quadform_energy(verts, tets, order, youngs, poisson, density, U, lam, g, batch=4096):
    """sum_i g_i (u_i^T K u_i - lam_i u_i^T M u_i), differentiable in verts.
    This is the scalar whose vertex-gradient torch autograd produces when the
    reference back-propagates sum_i g_i * get_vals()[i] (diff_model.py:390-399):
    the lam_i in `predict` are constants and U is detached."""
    mu, la = lame(youngs, poisson)
    Ut = torch.as_tensor(U, dtype=torch.float64)
    lamt = torch.as_tensor(lam, dtype=torch.float64)
    gt = torch.as_tensor(g, dtype=torch.float64)
    total = torch.zeros((), dtype=torch.float64)
    for s in range(0, tets.shape[0], batch):
        sl = slice(s, min(s + batch, tets.shape[0]))
        d = element_dofs(tets[sl])
        ue = Ut[d]                                              # (T, m, k)
        Ke = element_stiffness(verts, tets, order, mu, la, sl)
        Me = element_mass(verts, tets, order, density, sl)
        qk = torch.einsum("tak,tab,tbk->k", ue, Ke, ue)
        qm = torch.einsum("tak,tab,tbk->k", ue, Me, ue)
        total = total + (gt * (qk - lamt * qm)).sum()
    return total


def eigval_grad_shape(verts, tets, order, youngs, poisson, density, U, lam, g):
    """d/d(verts) of sum_i g_i * get_vals()[i]  -> (V,3) fp32 like the reference."""
    v = verts.detach().clone().requires_grad_(True)
    e = quadform_energy(v, tets, order, youngs, poisson, density, U, lam, g)
    e.backward()
    return v.grad


def material_quadforms(verts, tets, order, U, batch=4096):
    """q_mu[i] = u_i^T K(mu=1,lam=0) u_i,  q_lam[i] = u_i^T K(mu=0,lam=1) u_i
    (the reference evaluates the same thing matrix-free in fp32 via
    stiff_func, diff_model.py:314-328 + deform.py:70-87,149-165)."""
    Ut = torch.as_tensor(U, dtype=torch.float64)
    k = Ut.shape[1]
    qmu = torch.zeros(k, dtype=torch.float64)
    qla = torch.zeros(k, dtype=torch.float64)
    for s in range(0, tets.shape[0], batch):
        sl = slice(s, min(s + batch, tets.shape[0]))
        ue = Ut[element_dofs(tets[sl])]
        qmu += torch.einsum("tak,tab,tbk->k", ue, element_stiffness(verts, tets, order, 1.0, 0.0, sl), ue)
        qla += torch.einsum("tak,tab,tbk->k", ue, element_stiffness(verts, tets, order, 0.0, 1.0, sl), ue)
    return qmu.numpy(), qla.numpy()


def undamped_freqs(lam):
    """f = sqrt(lambda) / 2pi (diff_model.py:387)."""
    return np.sqrt(lam) / 2 / np.pi


# ----------------------------------------------------------------------------
# modal synthesis
# ----------------------------------------------------------------------------
def synth_closed_form(amp, damp, freq_d, T, sr, dtype=np.float64):
    """y[b,t] = sum_m a[b,m] exp(-d[m](t+1)/sr) sin(2 pi fd[m] (t+1)/sr): the
    exact value of the cumsum formulation (oscillator.py:297-304).
    amp (B,k), damp (k,) or (B,k), freq_d same."""
    amp = np.asarray(amp, dtype)
    damp = np.broadcast_to(np.asarray(damp, dtype), amp.shape)
    fd = np.broadcast_to(np.asarray(freq_d, dtype), amp.shape)
    tau = (np.arange(T, dtype=dtype) + 1) / sr
    y = np.zeros((amp.shape[0], T), dtype)
    for m in range(amp.shape[1]):
        y += amp[:, m:m + 1] * np.exp(-damp[:, m:m + 1] * tau) * np.sin(2 * np.pi * fd[:, m:m + 1] * tau)
    return y


def rayleigh_damping(freq, alpha, beta):
    """lambda=(2 pi f)^2; d = (alpha + beta lambda)/2; f_d = sqrt(lambda - d^2)/2pi
    (oscillator.py:287-292)."""
    lbd = (np.asarray(freq, np.float64) * 2 * np.pi) ** 2
    d = 0.5 * (alpha + beta * lbd)
    return d, np.sqrt(lbd - d ** 2) / (2 * np.pi)


def apply_force(signal, force):
    """Causal FIR with the force, truncated to T samples: the flipped-kernel
    grouped conv1d with padding F-1 (oscillator.py:305-309)."""
    T = signal.shape[1]
    return np.stack([np.convolve(signal[b], force[b])[:T] for b in range(signal.shape[0])])


def traditional_oscillator(freq, alpha, beta, force, T, sr):
    """TraditionalDampedOscillator.forward in fp64 (oscillator.py:282-310)."""
    B = force.shape[0]
    d, fd = rayleigh_damping(np.asarray(freq).reshape(-1), alpha, beta)
    amp = np.ones((B, d.shape[0]))
    return apply_force(synth_closed_form(amp, d, fd, T, sr), np.asarray(force, np.float64))


# ----------------------------------------------------------------------------
# synthetic meshes
# ----------------------------------------------------------------------------
def kuhn_cube(N):
    """(N+1)^3 grid on [0,1]^3, vertex id (i(N+1)+j)(N+1)+k, six Kuhn tets per
    cell, permutation-major ordering."""
    lin = torch.linspace(0, 1, N + 1)
    gx, gy, gz = torch.meshgrid(lin, lin, lin, indexing="ij")
    verts = torch.stack([gx, gy, gz], dim=-1).reshape(-1, 3).to(torch.float32)
    ii, jj, kk = torch.meshgrid(torch.arange(N), torch.arange(N), torch.arange(N), indexing="ij")
    base = torch.stack([ii, jj, kk], dim=-1).reshape(-1, 3)
    tets = []
    for perm in itertools.permutations(range(3)):
        p = base.clone()
        ids = [(p[:, 0] * (N + 1) + p[:, 1]) * (N + 1) + p[:, 2]]
        for ax in perm:
            p = p.clone()
            p[:, ax] += 1
            ids.append((p[:, 0] * (N + 1) + p[:, 1]) * (N + 1) + p[:, 2])
        tets.append(torch.stack(ids, dim=1))
    return verts, torch.cat(tets, dim=0).long()
