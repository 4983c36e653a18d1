"""GPU parity tests of the individual kernels behind the C-ABI against the CPU oracle.

Tolerances:
  pattern           bit-exact
  K values          <= 2e-6 * max|K|   (fp32-geometry floor: the kernel inverts A in fp64,
                                         the reference in fp32)
  M values          <= 1e-12 relative
  SpMM / Gram / GEMM  <= 1e-12 relative to the fp64 result norm
"""
import hashlib

import numpy as np
import pytest
import scipy.linalg
import torch

from conftest import golden
from oracle import modal_oracle as mo

pytestmark = pytest.mark.gpu

STEEL = (7850.0, 2.0e11, 0.29)


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _setup(meshes, name, order):
    from diffsound_b200 import native
    from diffsound_b200.diffelastic import mass_matrix as mmx
    v, t = meshes[name]
    pv, pt = mo.promote(torch.tensor(v), torch.tensor(t), order)
    dev = torch.device("cuda:0")
    verts = pv.to(dev).contiguous()
    tets = pt.to(torch.int32).to(dev).contiguous()
    pat = native.Pattern(tets, verts.shape[0])
    mu, lam = mo.lame(STEEL[1], STEEL[2])
    ctab = mmx.stiffness_contraction_table(order).to(dev)
    mtab = mmx.mass_density_table(order, STEEL[0]).to(dev)
    Kval, Mblk = native.assemble_km(verts, tets, order, pat, mu, lam, ctab, mtab)
    return pv, pt, pat, Kval, Mblk


@pytest.mark.parametrize("name,order", [("cube2", 1), ("cube2", 2), ("cube3", 1), ("cube3", 2),
                                        ("grid16", 1), ("grid16", 2), ("bowl", 1), ("bowl", 2)])
def test_pattern_bit_exact(meshes, name, order):
    pv, pt, pat, _, _ = _setup(meshes, name, order)
    crow, col, brow, bcol = mo.pattern(pt, pv.shape[0])
    c, k = pat.csr()
    assert c.dtype == torch.int64 and k.dtype == torch.int64
    assert np.array_equal(c.cpu().numpy(), crow)
    assert np.array_equal(k.cpu().numpy(), col)
    assert np.array_equal(pat.brow.cpu().numpy(), brow)
    assert np.array_equal(pat.bcol.cpu().numpy(), bcol)
    g = golden(f"modal_{name}_o{order}")
    assert _sha(c.cpu().numpy()) == str(g["crow_sha"])
    assert _sha(k.cpu().numpy()) == str(g["col_sha"])
    assert int(g["nnz"]) == pat.nnz
    # contributor lists: every element entry appears exactly once, grouped by slot
    contrib = pat.contrib.cpu().numpy()
    assert np.array_equal(np.sort(contrib), np.arange(contrib.size))


@pytest.mark.parametrize("name", ["cube2", "cube3", "grid16", "bowl"])
def test_assembly_tet_sequential_matches_row_kernel(meshes, name):
    """The two assembly kernels of quadratic meshes (tet-sequential rows = the default, balanced contributor lists = the
    fall-back for rows of more than 256 blocks) sum the same element contributions in the same ascending-tet order per
    slot; they differ only in how a slot's sum is associated (<= a few ulp of the largest contribution)."""
    from diffsound_b200 import native
    from diffsound_b200.diffelastic import mass_matrix as mmx
    v, t = meshes[name]
    pv, pt = mo.promote(torch.tensor(v), torch.tensor(t), 2)
    dev = torch.device("cuda:0")
    verts, tets = pv.to(dev).contiguous(), pt.to(torch.int32).to(dev).contiguous()
    pat = native.Pattern(tets, verts.shape[0])
    assert pat.slot is not None and 1 <= pat.max_deg <= 256
    mu, lam = mo.lame(STEEL[1], STEEL[2])
    ctab, mtab = mmx.stiffness_contraction_table(2).to(dev), mmx.mass_density_table(2, STEEL[0]).to(dev)
    K1, M1 = native.assemble_km(verts, tets, 2, pat, mu, lam, ctab, mtab, kernel="tets")
    K2, M2 = native.assemble_km(verts, tets, 2, pat, mu, lam, ctab, mtab, kernel="rows")
    assert float((K1 - K2).abs().max()) <= 1e-13 * float(K2.abs().max())
    assert float((M1 - M2).abs().max()) <= 1e-14 * float(M2.abs().max())
    # a second call reuses the cleared row images: identical bits
    K3, M3 = native.assemble_km(verts, tets, 2, pat, mu, lam, ctab, mtab, kernel="tets")
    assert torch.equal(K1, K3) and torch.equal(M1, M3)


@pytest.mark.parametrize("name,order", [("cube3", 1), ("cube3", 2), ("grid16", 1), ("grid16", 2), ("bowl", 2)])
def test_assembly_vs_oracle_and_golden(meshes, name, order):
    from diffsound_b200 import native
    pv, pt, pat, Kval, Mblk = _setup(meshes, name, order)
    K, M = mo.assemble(pv, pt, order, STEEL[1], STEEL[2], STEEL[0])
    kv = Kval.cpu().numpy()
    mv = native.mass_expand(pat, Mblk).cpu().numpy()
    assert np.abs(kv - K.data).max() <= 2e-6 * np.abs(K.data).max()
    nz = M.data != 0
    assert np.array_equal(mv == 0, ~nz)
    assert (np.abs(mv[nz] - M.data[nz]) / np.abs(M.data[nz])).max() <= 1e-12
    g = golden(f"modal_{name}_o{order}")
    if tuple(g["material"][:3]) == STEEL:
        idx = g["sample_idx"]
        assert np.abs(kv[idx] - g["K_sample"]).max() <= 2e-6 * float(g["K_absmax"])
        assert np.allclose(mv[idx], g["M_sample"], rtol=1e-12, atol=0)
    # rigid translations are in the null space of K (to fp32-geometry precision)
    n = K.shape[0]
    tr = np.zeros((n, 3))
    for c in range(3):
        tr[c::3, c] = 1
    assert np.abs(K @ tr).max() <= 1e-5 * np.abs(K.data).max()


@pytest.mark.parametrize("ncols", [16, 32, 48, 64])
def test_spmm_vs_scipy(meshes, ncols):
    from diffsound_b200 import native
    pv, pt, pat, Kval, Mblk = _setup(meshes, "grid16", 2)
    K, M = mo.assemble(pv, pt, 2, STEEL[1], STEEL[2], STEEL[0])
    Kc = K.copy(); Kc.data = Kval.cpu().numpy()
    Mc = M.copy(); Mc.data = native.mass_expand(pat, Mblk).cpu().numpy()
    rng = np.random.default_rng(0)
    X = rng.standard_normal((K.shape[0], ncols))
    Xd = torch.tensor(X, device="cuda:0")
    Y = native.spmm(pat, Kval, None, Xd).cpu().numpy()
    ref = Kc @ X
    assert np.abs(Y - ref).max() <= 1e-12 * np.abs(ref).max()
    Y = native.spmm(pat, None, Mblk, Xd, shift=1.0).cpu().numpy()
    ref = Mc @ X
    assert np.abs(Y - ref).max() <= 1e-12 * np.abs(ref).max()
    Y0 = torch.tensor(rng.standard_normal(X.shape), device="cuda:0")
    Y = native.spmm(pat, Kval, Mblk, Xd, shift=3.0e4, alpha=-0.5, beta=2.0, Y0=Y0).cpu().numpy()
    ref = -0.5 * (Kc @ X + 3.0e4 * (Mc @ X)) + 2.0 * Y0.cpu().numpy()
    assert np.abs(Y - ref).max() <= 1e-12 * np.abs(ref).max()
    YK, YM = native.spmm_k_and_m(pat, Kval, Mblk, Xd)
    assert np.abs(YK.cpu().numpy() - Kc @ X).max() <= 1e-12 * np.abs(Kc @ X).max()
    assert np.abs(YM.cpu().numpy() - Mc @ X).max() <= 1e-12 * np.abs(Mc @ X).max()
    # strided views (a column block of a wider buffer)
    wide = torch.zeros(K.shape[0], 96, dtype=torch.float64, device="cuda:0")
    wide[:, 16:16 + ncols] = Xd[:, :min(ncols, 80)]
    if 16 + ncols <= 96:
        out = torch.zeros_like(wide)
        native.spmm(pat, Kval, None, wide[:, 16:16 + ncols], out=out[:, 16:16 + ncols])
        assert np.abs(out[:, 16:16 + ncols].cpu().numpy() - Kc @ X).max() <= 1e-12 * np.abs(Kc @ X).max()
        assert float(out[:, :16].abs().max()) == 0.0


@pytest.mark.parametrize("n,p,q", [(1000, 16, 16), (4097, 48, 48), (30011, 64, 32), (7, 8, 64), (100000, 48, 16)])
def test_gram_dmma(n, p, q):
    from diffsound_b200 import native
    g = torch.Generator(device="cuda:0").manual_seed(1)
    A = torch.randn(n, p, dtype=torch.float64, device="cuda:0", generator=g)
    B = torch.randn(n, q, dtype=torch.float64, device="cuda:0", generator=g)
    G = native.gram(A, B)
    ref = A.T @ B
    assert float((G - ref).abs().max()) <= 1e-12 * float(ref.abs().max()) * np.sqrt(n)
    # sub-blocks of a wider buffer
    wide = torch.randn(n, 96, dtype=torch.float64, device="cuda:0", generator=g)
    G = native.gram(wide[:, 16:16 + p], wide[:, 32:32 + q])
    ref = wide[:, 16:16 + p].T @ wide[:, 32:32 + q]
    assert float((G - ref).abs().max()) <= 1e-12 * float(ref.abs().max()) * np.sqrt(n)


@pytest.mark.parametrize("n,m,prow,q2", [(1003, 48, 144, 48), (777, 48, 96, 32), (4099, 32, 96, 16), (515, 16, 48, 16),
                                         (100000, 48, 144, 16)])
def test_rr_update_fused(n, m, prow, q2):
    """Fused LOBPCG basis update against torch matmul: X' = [X W P] C1, P' = [W P] C2 for S, KS, MS at once."""
    from diffsound_b200 import native
    DEV = "cuda:0"
    g = torch.Generator(device=DEV).manual_seed(5)
    bufs = [torch.randn(n, 3 * m, dtype=torch.float64, device=DEV, generator=g) for _ in range(3)]
    Cfull = torch.randn(144, 144, dtype=torch.float64, device=DEV, generator=g)
    C2full = torch.randn(144, 144, dtype=torch.float64, device=DEV, generator=g)
    C1, C2 = Cfull[:prow, :m], C2full[:prow - m, :q2]
    outs = native.rr_update(bufs, prow, m, C1, C2, q2)
    for A, Y in zip(bufs, outs):
        ref1 = A[:, :prow] @ C1
        ref2 = A[:, m:prow] @ C2
        assert torch.allclose(Y[:, :m], ref1, rtol=1e-12, atol=1e-11)
        assert torch.allclose(Y[:, 2 * m:2 * m + q2], ref2, rtol=1e-12, atol=1e-11)
        assert float(Y[:, m:2 * m].abs().max()) == 0.0


@pytest.mark.parametrize("n,p,q", [(1000, 16, 16), (4099, 48, 48), (30011, 144, 48), (5, 4, 8), (50000, 32, 64)])
def test_block_gemm_dmma(n, p, q):
    from diffsound_b200 import native
    g = torch.Generator(device="cuda:0").manual_seed(2)
    A = torch.randn(n, p, dtype=torch.float64, device="cuda:0", generator=g)
    Cm = torch.randn(p, q, dtype=torch.float64, device="cuda:0", generator=g)
    Y = native.block_gemm(A, Cm)
    ref = A @ Cm
    assert float((Y - ref).abs().max()) <= 1e-12 * float(ref.abs().max()) * np.sqrt(p)
    Y0 = torch.randn(n, q, dtype=torch.float64, device="cuda:0", generator=g)
    Y = Y0.clone()
    native.block_gemm(A, Cm, beta=-0.25, out=Y)
    ref = A @ Cm - 0.25 * Y0
    assert float((Y - ref).abs().max()) <= 1e-12 * float(ref.abs().max()) * np.sqrt(p)


@pytest.mark.parametrize("N", [6, 7, 33, 48, 95, 96, 114, 131, 143, 144])
def test_eigh_generalized_jacobi(N):
    from diffsound_b200 import native
    rng = np.random.default_rng(N)
    Q = rng.standard_normal((4 * N, N))
    GM = Q.T @ Q / (4 * N) + 0.1 * np.eye(N)
    ev = np.concatenate([np.abs(rng.standard_normal(6)) * 1e-3, 10 ** rng.uniform(7, 12, N - 6)]) if N > 6 \
        else 10 ** rng.uniform(7, 9, N)
    V = np.linalg.qr(rng.standard_normal((N, N)))[0]
    Lm = np.linalg.cholesky(GM)
    GK = Lm @ (V * ev) @ V.T @ Lm.T
    GK = (GK + GK.T) / 2
    sigma = 1e5
    theta, Cm, info = native.eigh_generalized(torch.tensor(GK, device="cuda:0"), torch.tensor(GM, device="cuda:0"), sigma)
    info = info.cpu().numpy()
    assert info[0] == 0, info
    w, Z = scipy.linalg.eigh(GK, GM)
    th = theta.cpu().numpy()
    assert np.all(np.diff(th) >= 0)
    assert np.abs(th - w).max() <= 1e-11 * (np.abs(w).max() + sigma) or np.allclose(th, w, rtol=1e-9, atol=1e-9 * sigma)
    Cn = Cm.cpu().numpy()
    assert np.abs(Cn.T @ GM @ Cn - np.eye(N)).max() <= 1e-10
    big = w > 1e3
    R = GK @ Cn - GM @ Cn * th
    assert np.abs(R[:, big]).max() <= 1e-9 * np.abs(GK).max()


def test_legacy_mass_coo_matches_reference_layout(meshes):
    """Same call signature and output layout as the reference's assemble_mass_matrix
    (src/cuda/massMatrixDouble.cu:138-158), checked against the oracle element matrices."""
    from diffsound_b200 import native
    from diffsound_b200.diffelastic.mass_matrix import get_elememt_mass_matrix
    for order in (1, 2):
        v, t = meshes["cube3"]
        pv, pt = mo.promote(torch.tensor(v), torch.tensor(t), order)
        dev = "cuda:0"
        vnum = pt.shape[1]
        msz = 3 * vnum
        T = pt.shape[0]
        vertices = pv.double().reshape(-1).to(dev)
        tets = pt.to(torch.int32).reshape(-1).to(dev)
        emm = get_elememt_mass_matrix(order).double().to(dev)
        values = torch.zeros(msz * msz * T, dtype=torch.float64, device=dev)
        rows = torch.zeros(msz * msz * T, dtype=torch.int32, device=dev)
        cols = torch.zeros_like(rows)
        native.assemble_mass_coo(vertices, tets, values, rows, cols, emm, 2700.0, order)
        Me = mo.element_mass(pv, pt, order, 1.0).numpy() * 2700.0      # (T, msz, msz)
        d = mo.element_dofs(pt).numpy()
        assert np.allclose(values.cpu().numpy().reshape(T, msz, msz), Me, rtol=1e-6, atol=0)
        assert np.array_equal(rows.cpu().numpy().reshape(T, msz, msz), np.broadcast_to(d[:, :, None], (T, msz, msz)))
        assert np.array_equal(cols.cpu().numpy().reshape(T, msz, msz), np.broadcast_to(d[:, None, :], (T, msz, msz)))


@pytest.mark.parametrize("ld,tiles", [(144, list(range(18))), (144, [0, 1, 2, 3, 4, 5, 6, 7, 8, 12, 13]), (96, list(range(12))),
                                      (48, list(range(6)))])
def test_gram_sym2_matches_fp64(ld, tiles):
    """Fused GK = S^T KS, GM = S^T MS (upper-triangle tiles over the active tile columns) against
    torch fp64 matmul: <= 1e-12 of the largest entry; inactive tiles stay untouched."""
    import ctypes as C
    from diffsound_b200 import _lib, native
    lib = _lib.load()
    dev = torch.device("cuda:0")
    n = 50_003                     # not a multiple of the 16-row chunk
    g = torch.Generator(device=dev).manual_seed(3)
    S, KS, MS = (torch.randn(n, ld, dtype=torch.float64, device=dev, generator=g) for _ in range(3))
    GK = torch.full((144, 144), -7.0, dtype=torch.float64, device=dev)
    GM = torch.full((144, 144), -7.0, dtype=torch.float64, device=dev)
    part = torch.empty(lib.ds_gram_sym2_scratch_elems(), dtype=torch.float64, device=dev)
    arr = (C.c_int * len(tiles))(*tiles)
    _lib.check(lib.ds_gram_sym2_f64(native._p(S), native._p(KS), native._p(MS), ld, n, arr, len(tiles), native._p(GK),
                                    native._p(GM), 144, native._p(part), native._stream()), "ds_gram_sym2_f64")
    rk, rm = (S.T @ KS).cpu().numpy(), (S.T @ MS).cpu().numpy()
    gk, gm = GK.cpu().numpy(), GM.cpu().numpy()
    act = np.zeros(18, dtype=bool)
    act[tiles] = True
    for ti in range(18):
        for tj in range(18):
            blk = (slice(8 * ti, 8 * ti + 8), slice(8 * tj, 8 * tj + 8))
            if act[ti] and act[tj] and ti <= tj:
                assert np.abs(gk[blk] - rk[blk]).max() <= 1e-12 * np.abs(rk).max()
                assert np.abs(gm[blk] - rm[blk]).max() <= 1e-12 * np.abs(rm).max()
            else:
                assert (gk[blk] == -7.0).all() and (gm[blk] == -7.0).all()


DEV = "cuda:0"


def test_gram_strip_matches_torch():
    from diffsound_b200 import native
    torch.manual_seed(3)
    n = 5003
    for wa in (16, 32, 48):
        S = torch.randn(n, 144, dtype=torch.float64, device=DEV)
        KS = torch.randn(n, 144, dtype=torch.float64, device=DEV)
        MS = torch.randn(n, 144, dtype=torch.float64, device=DEV)
        GsK, GsM = native.gram_strip(KS[:, 48:48 + wa], MS[:, 48:48 + wa], S)
        rk, rm = KS[:, 48:48 + wa].T @ S, MS[:, 48:48 + wa].T @ S
        assert float((GsK - rk).abs().max()) <= 1e-11 * float(rk.abs().max())
        assert float((GsM - rm).abs().max()) <= 1e-11 * float(rm.abs().max())


def test_rr_update2_matches_torch():
    from diffsound_b200 import native
    torch.manual_seed(4)
    n = 4099
    for m, wa, use_p in ((48, 48, True), (48, 32, True), (48, 16, False), (32, 32, True), (16, 16, True)):
        bufs = [torch.randn(n, 3 * m, dtype=torch.float64, device=DEV) for _ in range(3)]
        Cm = torch.randn(3 * m, 3 * m, dtype=torch.float64, device=DEV)
        Cm[m + wa:2 * m] = 0
        if not use_p:
            Cm[2 * m:] = 0
        outs = native.rr_update2(bufs, m, wa, use_p, Cm)
        for A, Y in zip(bufs, outs):
            P = A[:, m:] @ Cm[m:, :m]
            X = A[:, :m] @ Cm[:m, :m] + P
            sc = float(X.abs().max())
            assert float((Y[:, 2 * m:] - P).abs().max()) <= 1e-12 * sc
            assert float((Y[:, :m] - X).abs().max()) <= 1e-12 * sc


def test_gram_algebra_matches_torch():
    from diffsound_b200 import native
    torch.manual_seed(5)
    for m in (16, 48):
        N = 3 * m
        B = torch.randn(N, N, dtype=torch.float64, device=DEV)
        GK, GM = B @ B.T, B.T @ B + torch.eye(N, dtype=torch.float64, device=DEV)
        Cm = torch.randn(N, N, dtype=torch.float64, device=DEV)
        theta = torch.rand(N, dtype=torch.float64, device=DEV)
        GKn, GMn = native.gram_algebra(GK, GM, Cm, theta, m)
        C1 = Cm[:, :m]
        Cwp = C1.clone()
        Cwp[:m] = 0
        for G, Gn, diag in ((GK, GKn, theta[:m]), (GM, GMn, torch.ones(m, dtype=torch.float64, device=DEV))):
            sc = float((C1.T @ G @ C1).abs().max())
            assert float((Gn[:m, :m] - torch.diag(diag)).abs().max()) == 0.0
            assert float((Gn[:m, 2 * m:] - C1.T @ G @ Cwp).abs().max()) <= 1e-12 * sc
            assert float((Gn[2 * m:, :m] - (C1.T @ G @ Cwp).T).abs().max()) <= 1e-12 * sc
            assert float((Gn[2 * m:, 2 * m:] - Cwp.T @ G @ Cwp).abs().max()) <= 1e-12 * sc
            assert float(Gn[m:2 * m].abs().max()) == 0.0 and float(Gn[:, m:2 * m].abs().max()) == 0.0


def test_fp64_peak_runs():
    from diffsound_b200 import native
    assert native.fp64_peak(0, 256, 1) > 1.0 and native.fp64_peak(1, 256, 1) > 1.0
