"""`lobpcg` / `lobpcg_func`: generalized symmetric eigenproblem A x = lambda B x, lowest pairs.

API mirror of src/lobpcg/_lobpcg.py:8-212 (same argument names and order, same
return convention: `E[:k]`, `X[:, :k]` [, rerr], ascending, B-orthonormal
columns).  The worker is the device-resident LOBPCG of csrc/lobpcg.cu
(`ds_lobpcg`), not a torch re-implementation.

Two drivers sit behind the one signature:

* operators with the FEM structure this library is built around -- A symmetric with dense
  3x3 node blocks, B = (one scalar per block) (x) I3 on (a subset of) the same block
  pattern, which is exactly what the reference's K and M are (SURVEY.md section 8:
  "nnz = 9 x node pairs") -- asked for their LOWEST pairs without a user preconditioner or
  tracker go to the device-resident solver `ds_lobpcg` (its own FP32 Chebyshev / two-level
  preconditioner);
* everything else the reference's signature allows -- a callable `A` (the reason
  `lobpcg_func` exists, _lobpcg.py:123-212), dense or unstructured sparse tensors,
  `largest=True` (the reference's default), a preconditioner `iK` (tensor or callable,
  _lobpcg.py:453,475), `tracker` / `profiler` hooks (_lobpcg.py:350-376) -- goes to the
  operator-form worker `_generic.LOBPCG`, whose dense steps run on the library's FP64
  Gram / eigh / block-GEMM kernels.

Computation is FP64 regardless of the input dtype (the reference's copy is FP32-only because
of its dtype table); results are cast back.  CPU tensors raise: there is no
CPU path.  Blocks wider than 48 columns raise (Rayleigh-Ritz limit 3 x 48 = 144).
"""
from typing import Optional

import torch

from .. import native


class BlockMatrices:
    """K (scalar CSR values on the block pattern) and M (one scalar per block)."""

    def __init__(self, pattern, Kval, Mblk):
        self.pattern, self.Kval, self.Mblk = pattern, Kval, Mblk


class _RawPattern:
    """A block-CSR pattern given directly by (brow, bcol), e.g. derived from torch sparse tensors."""

    def __init__(self, brow, bcol):
        self.brow, self.bcol = brow, bcol
        self.n_nodes = brow.numel() - 1
        self.nnzb = bcol.numel()
        self.device = brow.device

    @property
    def n(self):
        return 3 * self.n_nodes


def _coo_parts(A):
    if A.layout == torch.sparse_csr:
        A = A.to_sparse_coo()
    if A.layout != torch.sparse_coo:
        raise TypeError("lobpcg: A and B must be torch sparse (COO or CSR) tensors")
    A = A.coalesce()
    return A.indices(), A.values()


def from_torch_sparse(A, B=None):
    """Re-pack torch sparse A (and B) into the library's block layout.  Raises ValueError if the
    matrices do not have the 3x3-block / scalar-block structure."""
    if not A.is_cuda:
        raise RuntimeError("lobpcg: operators must be CUDA tensors (there is no CPU path)")
    n = A.shape[-1]
    if A.shape[-2] != n or n % 3 != 0:
        raise ValueError(f"lobpcg: A must be square with 3 dofs per node (got {tuple(A.shape)})")
    nb = n // 3
    ia, va = _coo_parts(A)
    keys = (ia[0] // 3) * nb + (ia[1] // 3)
    if B is not None:
        ib, vb = _coo_parts(B)
        kb = (ib[0] // 3) * nb + (ib[1] // 3)
        allk = torch.cat([keys, kb])
    else:
        allk = torch.cat([keys, torch.arange(nb, device=A.device) * (nb + 1)])
    uk = torch.unique(allk)                           # sorted block keys
    bi, bj = uk // nb, uk % nb
    deg = torch.bincount(bi, minlength=nb)
    brow = torch.zeros(nb + 1, dtype=torch.int64, device=A.device)
    brow[1:] = torch.cumsum(deg, 0)
    slot = torch.searchsorted(uk, keys)
    p = slot - brow[ia[0] // 3]
    off = 9 * brow[ia[0] // 3] + (ia[0] % 3) * 3 * deg[ia[0] // 3] + 3 * p + (ia[1] % 3)
    Kval = torch.zeros(9 * uk.numel(), dtype=torch.float64, device=A.device)
    Kval[off] = va.double()
    Mblk = torch.zeros(uk.numel(), dtype=torch.float64, device=A.device)
    if B is None:
        Mblk[torch.searchsorted(uk, torch.arange(nb, device=A.device) * (nb + 1))] = 1.0
    else:
        offd = (ib[0] % 3) != (ib[1] % 3)
        if bool((vb[offd] != 0).any()):
            raise ValueError("lobpcg: B must be (scalar per node block) (x) I3; found non-zero off-diagonal block entries")
        sb = torch.searchsorted(uk, kb)
        d0 = torch.zeros(uk.numel(), 3, dtype=torch.float64, device=A.device)
        sel = ~offd
        d0[sb[sel], (ib[0] % 3)[sel]] = vb[sel].double()
        if bool(((d0 - d0[:, :1]).abs() > 1e-12 * d0.abs().max()).any()):
            raise ValueError("lobpcg: B must be (scalar per node block) (x) I3; block diagonals differ")
        Mblk = d0[:, 0].contiguous()
    pat = _RawPattern(brow.to(torch.int32).contiguous(), bj.to(torch.int32).contiguous())
    return BlockMatrices(pat, Kval, Mblk)


def _resolve(A, B):
    if isinstance(A, BlockMatrices):
        return A
    owner = getattr(A, "__self__", None)
    if callable(A) and owner is not None and hasattr(owner, "_Kval") and owner._Kval is not None:
        return BlockMatrices(owner.deform.pattern, owner._Kval, owner._Mblk)     # DiffSoundObj.stiff_func
    if callable(A) and not torch.is_tensor(A):
        raise TypeError("matrix-free callable: operator-form driver")
    if not torch.is_tensor(A) or A.layout == torch.strided or (B is not None and (not torch.is_tensor(B) or B.layout == torch.strided)):
        raise TypeError("dense or non-tensor operator: operator-form driver")
    return from_torch_sparse(A, B)


def _generic(A, k, B, X, E, n, iK, niter, tol, largest, method, tracker, ortho_iparams, ortho_fparams, ortho_bparams,
             return_rerr, profiler):
    """Operator-form driver: mirrors the parameter handling of _lobpcg.py:27-121 / :139-212."""
    from ._generic import LOBPCG
    ref = next((t for t in (X, B, A, iK) if torch.is_tensor(t)), None)
    if ref is None:
        raise TypeError("lobpcg: at least one of A, B, X, iK must be a tensor (to know the size and the device)")
    if not ref.is_cuda:
        raise RuntimeError("lobpcg: operators must be CUDA tensors (there is no CPU path)")
    if torch.is_tensor(A):
        assert A.shape[-2] == A.shape[-1], A.shape
        if torch.is_tensor(B):
            assert A.shape == B.shape, (A.shape, B.shape)
    msize = (A if torch.is_tensor(A) else B if torch.is_tensor(B) else X).shape[-2]
    dtype = X.dtype if X is not None else (ref.dtype if ref.dtype in (torch.float32, torch.float64) else torch.float32)
    if tol is None:
        tol = {torch.float32: 1.2e-07, torch.float64: 2.23e-16}[dtype] ** 0.5
    k = (1 if X is None else X.shape[-1]) if k is None else k
    n = (k if n is None else n) if X is None else X.shape[-1]
    if msize < 3 * n:
        raise ValueError("LPBPCG algorithm is not applicable when the number of A rows (={})"
                         " is smaller than 3 x the number of requested eigenpairs (={})".format(msize, n))
    method = "ortho" if method is None else method
    iparams = {"m": msize, "n": n, "k": k, "niter": 1000 if niter is None else niter}
    fparams = {"tol": tol}
    bparams = {"largest": True if largest is None else largest}
    for extra, tgt in ((ortho_iparams, iparams), (ortho_fparams, fparams), (ortho_bparams, bparams)):
        if method == "ortho" and extra is not None:
            tgt.update(extra)
    if X is None:
        X = torch.randn((msize, n), dtype=dtype, device=ref.device)
    assert len(X.shape) == 2 and X.shape == (msize, n), (X.shape, (msize, n))
    worker = LOBPCG(A, B, X, E, iK, iparams, fparams, bparams, method, tracker, profiler)
    worker.run()
    if return_rerr:
        return worker.E[:k], worker.X[:, :k], worker.tvars["rerr"]
    return worker.E[:k], worker.X[:, :k]


def lobpcg(A, k: Optional[int] = None, B=None, X=None, E=None, n: Optional[int] = None, iK=None,
           niter: Optional[int] = None, tol: Optional[float] = None, largest: Optional[bool] = None,
           method: Optional[str] = None, tracker=None, ortho_iparams=None, ortho_fparams=None, ortho_bparams=None,
           return_rerr=False, profiler=None):
    largest = True if largest is None else largest
    bm = None
    if not largest and iK is None and tracker is None and profiler is None:
        try:
            bm = _resolve(A, B)
        except (TypeError, ValueError):
            bm = None                    # not the FEM block structure: operator-form driver
    if bm is None:
        return _generic(A, k, B, X, E, n, iK, niter, tol, largest, method, tracker, ortho_iparams, ortho_fparams,
                        ortho_bparams, return_rerr, profiler)
    pat = bm.pattern
    msize = pat.n
    k = (1 if X is None else X.shape[-1]) if k is None else k
    n = (k if n is None else n) if X is None else X.shape[-1]
    if msize < 3 * n:
        raise ValueError("LPBPCG algorithm is not applicable when the number of A rows (={})"
                         " is smaller than 3 x the number of requested eigenpairs (={})".format(msize, n))
    out_dtype = X.dtype if X is not None else (A.dtype if torch.is_tensor(A) and A.dtype.is_floating_point else torch.float64)
    mcols = next((c for c in (16, 32, 48) if c >= n), None)
    if mcols is None:
        raise NotImplementedError(f"lobpcg: block size {n} exceeds the solver's 48 columns")
    gen = torch.Generator(device=pat.device).manual_seed(0)
    Xw = torch.randn(msize, mcols, dtype=torch.float64, device=pat.device, generator=gen)
    if X is not None:
        assert X.shape == (msize, n), (X.shape, (msize, n))
        Xw[:, :n] = X.double()
    tol = 1e-6 if tol is None else max(float(tol), 1e-12)
    niter = 1000 if niter is None else niter
    deg = int(min(40, max(8, round(msize ** (1.0 / 3.0) / 3.0))))
    lam, res, stats = native.lobpcg(pat, bm.Kval, bm.Mblk, Xw, nev=k, tol=tol, maxit=niter, cheb_degree=deg,
                                    cheb_ratio=0.4 * deg * deg, n_rigid=-1)
    Eo = lam[:k].to(out_dtype)
    Xo = Xw[:, :k].to(out_dtype)
    if return_rerr:
        return Eo, Xo, res[:k].to(out_dtype)
    return Eo, Xo


def lobpcg_func(A, B, k: Optional[int] = None, X=None, E=None, n: Optional[int] = None, iK=None,
                niter: Optional[int] = None, tol: Optional[float] = None, largest: Optional[bool] = None,
                method: Optional[str] = None, tracker=None, ortho_iparams=None, ortho_fparams=None,
                ortho_bparams=None, return_rerr=False, profiler=None):
    """Same as `lobpcg` with the reference's argument order (A may be a callable there,
    _lobpcg.py:123-212); see `_resolve` for what is accepted here."""
    return lobpcg(A, k=k, B=B, X=X, E=E, n=n, iK=iK, niter=niter, tol=tol, largest=largest, method=method,
                  tracker=tracker, ortho_iparams=ortho_iparams, ortho_fparams=ortho_fparams,
                  ortho_bparams=ortho_bparams, return_rerr=return_rerr, profiler=profiler)
