"""Differentiable modal model: assemble K, M -> lowest modes -> differentiable eigenvalues.

API mirror of src/diffelastic/diff_model.py:17-399 (`FixedLinear`,
`TrainableLinear`, `build_model`, `DiffSoundObj`): same constructor arguments,
attribute names (`tetmesh, deform, material_model, mode_num, task, stiff_matrix,
mass_matrix, eigenvalues, U_hat, U_hat_full`), method names, output shapes and
dtypes, so the reference's experiment loops run against this class unchanged.

Underneath, every step is one or a few calls into libdiffsound_sm100.so:

  reference (torch eager / SciPy)                      here
  update_mass_matrix + update_stiff_matrix  (:184-312)  ds_assemble_km into a fixed block-CSR pattern
  eigen_decomposition_arpack (CPU eigsh)    (:335-369)  ds_lobpcg (device-resident, preconditioned)
  get_vals (sparse mm + autograd graph)     (:390-399)  ds_spmm_k_and_m / ds_gram_f64 forward,
                                                        ds_eigval_grad_shape backward
  get_undamped_freqs (fp32 matrix-free K U) (:371-388)  ds_eigval_quadforms_material once per
                                                        decomposition; lambda_i(E, nu) = mu q_mu + lam q_lam
There is no CPU fallback: tensors must be CUDA tensors and the library must load.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import native
from ..ddsp.oscillator import WeightedParam
from .deform import Deform
from .mass_matrix import get_elememt_mass_matrix  # noqa: F401  (re-exported like the reference)
from .material_model import Material, MatSet
from .mesh import TetMesh


def _lame(youngs, poisson):
    lame_lambda = youngs * poisson / ((1 + poisson) * (1 - 2 * poisson))
    lame_mu = youngs / (2 * (1 + poisson))
    return lame_mu, lame_lambda


def _stress(F, lame_mu, lame_lambda):
    """P = mu (F + F^T) + lambda tr(F) I   (diff_model.py:39-41)."""
    tr = F.diagonal(dim1=-2, dim2=-1).sum(-1)
    return lame_mu * (F + F.transpose(1, 2)) + lame_lambda * tr[:, None, None] * torch.eye(3, device=F.device, dtype=F.dtype)


class FixedLinear(nn.Module):
    """Linear elasticity with fixed E, nu ("gt" and shape tasks)."""

    def __init__(self, mat: Material):
        super().__init__()
        self.youngs = mat.youngs
        self.poisson = mat.poisson
        self.mat = mat

    def lame(self):
        return _lame(self.youngs, self.poisson)

    def forward(self, F: torch.Tensor):
        b, n, _, _ = F.shape
        return self.get_stress(F.reshape(b * n, 3, 3)).reshape(b, n, 3, 3)

    def get_stress(self, F):
        mu, lam = self.lame()
        return _stress(F, mu, lam)

    def jacobian_F(self):
        """d(stress)/dF at F = 0 as a (1,3,3,1,3,3) fp64 tensor (diff_model.py:44-48); closed form."""
        mu, lam = self.lame()
        return _jacobian(float(mu), float(lam))


class TrainableLinear(nn.Module):
    """Learnable E, nu as softplus-normalised mixes of 16 bins (diff_model.py:51-96).  Like the
    reference the parameters live on the CPU (WeightedParam.values_list is not a buffer)."""

    def __init__(self, mat: Material, bin_num=16, baseline=False):
        super().__init__()
        self.youngs_list = torch.exp(torch.linspace(np.log(mat.youngs / 10), np.log(mat.youngs * 10), bin_num))
        if baseline:
            self.poisson_list = torch.linspace(mat.poisson, mat.poisson, 1)
        else:
            self.poisson_list = torch.linspace(0.01, 0.499, bin_num)
        self.youngs = WeightedParam(self.youngs_list)
        self.poisson = WeightedParam(self.poisson_list)
        self.mat = mat

    def lame(self):
        return _lame(self.youngs(), self.poisson())

    def forward(self, F: torch.Tensor):
        b, n, _, _ = F.shape
        return self.get_stress(F.reshape(b * n, 3, 3)).reshape(b, n, 3, 3)

    def get_stress(self, F):
        mu, lam = self.lame()
        return _stress(F, mu.to(F.device), lam.to(F.device))

    def jacobian_F(self):
        mu, lam = self.lame()
        return _jacobian(float(mu), float(lam))


def _jacobian(mu, lam):
    dev = torch.device("cuda", torch.cuda.current_device())
    eye = torch.eye(3, dtype=torch.float64, device=dev)
    J = (mu * (torch.einsum("ik,jl->ijkl", eye, eye) + torch.einsum("il,jk->ijkl", eye, eye))
         + lam * torch.einsum("ij,kl->ijkl", eye, eye))
    return J.reshape(1, 3, 3, 1, 3, 3)


def build_model(mesh_dir, mode_num, order, mat, task, vertices=None, tets=None, scale_range=None, init_scale=None):
    if task == "material" or task == "mat_baseline":
        mat_model = TrainableLinear
    elif task == "gt":
        mat_model = FixedLinear
    else:
        raise ValueError("task not defined")
    model = DiffSoundObj(mesh_dir=mesh_dir, mode_num=mode_num, order=order, mat=mat, mat_model=mat_model, task=task)
    if task == "material" or task == "mat_baseline":
        model.init_material_coeffs()
    return model


def _gram_diag(X, Y):
    """diag(X^T Y) for (n, k) fp64 blocks, in chunks the Gram kernel accepts (<= 64 columns, multiples of 8)."""
    k = X.shape[1]
    out = []
    for c0 in range(0, k, 48):
        c1 = min(k, c0 + 48)
        out.append(torch.diagonal(native.gram(X[:, c0:c1], Y[:, c0:c1])))
    return torch.cat(out)


class _EigvalShape(torch.autograd.Function):
    """get_vals(): value lambda + (u^T K u - lambda u^T M u), gradient u^T (dK - lambda dM) u w.r.t.
    the (promoted) vertex positions.  Everything the backward pass needs is pinned in ctx at forward
    time (eigenvectors, eigenvalues, geometry, material), as autograd does for the reference
    (diff_model.py:390-399): a later eigen_decomposition() / update_*_matrix() on the same object
    does not change the gradient of an earlier get_vals()."""

    @staticmethod
    def forward(ctx, vertices, obj):
        X, lam = obj._Xpad, obj.eigenvalues          # all wanted columns, padded to a multiple of 16
        d = obj.deform
        pat = d.pattern
        KX, MX = native.spmm_k_and_m(pat, obj._Kval, obj._Mblk, X)
        lo, hi = 6, 6 + obj.mode_num
        uku = _gram_diag(X, KX)[lo:hi]
        umu = _gram_diag(X, MX)[lo:hi]
        predict = torch.zeros(obj.mode_num, dtype=torch.float32, device=X.device)
        predict += lam
        predict += uku - lam * umu
        ctx.order = obj.tetmesh.order
        ctx.lame = obj._lame_used
        ctx.vdtype = obj.tetmesh.vertices.dtype
        inc_ptr, inc = d.incidence
        ctx.U = obj.U_hat                        # a view of storage that is never written again (see _start_block)
        ctx.save_for_backward(obj._verts32, d.tets_i32, d.ctab, d.mtab(obj._density_used), lam, inc_ptr, inc)
        return predict.unsqueeze(1)

    @staticmethod
    def backward(ctx, g):
        verts32, tets, ctab, mtab, lam, inc_ptr, inc = ctx.saved_tensors
        mu, lam_l = ctx.lame
        gv = g.reshape(-1).to(torch.float64).contiguous()
        grad = native.eigval_grad_shape(verts32, tets, ctx.order, mu, lam_l, ctab, mtab, ctx.U, lam, gv, inc_ptr, inc)
        return grad.to(ctx.vdtype), None


class _StiffFunc(torch.autograd.Function):
    """y = (mu K_mu + lam K_lam) x with gradients to x (K is symmetric: K^T g) and to the two Lame
    scalars (<g, K_mu x>, <g, K_lam x>) -- stiff_func of the reference is differentiable in both
    (diff_model.py:314-328 through deform.py:70-165)."""

    @staticmethod
    def forward(ctx, x, mu, lam, obj):
        Kmu, Kla = obj._unit_stiffness()
        pat = obj.deform.pattern
        cols = x.shape[1]
        ys = []
        for c0 in range(0, cols, 128):
            c1 = min(cols, c0 + 128)
            cp = (c1 - c0 + 15) // 16 * 16
            xp = torch.zeros(x.shape[0], cp, dtype=torch.float64, device=x.device)
            xp[:, :c1 - c0] = x[:, c0:c1]
            ys.append((native.spmm(pat, Kmu, None, xp)[:, :c1 - c0], native.spmm(pat, Kla, None, xp)[:, :c1 - c0]))
        ymu = torch.cat([a for a, _ in ys], dim=1)
        yla = torch.cat([b for _, b in ys], dim=1)
        ctx.obj = obj
        ctx.save_for_backward(ymu, yla, mu, lam)
        ctx.xdtype = x.dtype
        return (mu.double() * ymu + lam.double() * yla).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        ymu, yla, mu, lam = ctx.saved_tensors
        g64 = g.to(torch.float64)
        gx = gmu = gla = None
        if ctx.needs_input_grad[0]:
            gx = _StiffFunc.apply(g.contiguous(), mu.detach(), lam.detach(), ctx.obj).to(ctx.xdtype)
        if ctx.needs_input_grad[1]:
            gmu = (g64 * ymu).sum().to(mu.dtype)
        if ctx.needs_input_grad[2]:
            gla = (g64 * yla).sum().to(lam.dtype)
        return gx, gmu, gla, None


class DiffSoundObj:
    #: options of the eigensolver (see ds_lobpcg_opts in include/diffsound_sm100.h)
    # relative residual ||K u - lam M u|| / (lam ||M u||).  Eigenvalue error ~ tol^2, eigenvector (and
    # therefore d(lambda)/d(theta)) error ~ tol * lam / gap: 1e-5 keeps the gradient inside 1e-5.
    eig_tol = 1e-5
    two_level = True        # quadratic meshes: p-multigrid preconditioner (P1 coarse level) in the eigensolver
    nested_start = True     # ... and a P1 eigen-solve for the start block (skipped on a warm start)
    nested_tol = 3e-2
    nested_degree = 0       # 0: automatic
    smooth_steps = 3        # Chebyshev-Jacobi smoothing steps on the P2 operator before / after the coarse correction
    smooth_ratio = 8.0      # ... damping the upper [lmax / ratio, lmax] of the spectrum
    coarse_degree = 0       # Chebyshev steps of the P1 coarse solve (0: automatic, ~ n_coarse^(1/3) / 1.2)
    coarse_ratio = 0.0      # 0: automatic, 0.4 * degree^2
    cheb_degree = 0         # one-level Chebyshev preconditioner (linear meshes): polynomial degree, 0 = automatic
    morton = True           # the FP32 preconditioner keeps its operator in a Morton node numbering (SpMM locality)
    eig_maxit = 400
    # Sliver elements (marching-tets meshes) put stiff components into the residuals that exceed their smooth part by more than
    # FP32 resolves; the FP32 cycle then stalls.  Policy: meshes whose worst element is `sliver_ratio`
    # times flatter than the median go straight to the FP64 Chebyshev preconditioner (ds_lobpcg_opts.precond_fp64); others try
    # the FP32 cycle for at most `eig_fast_maxit` iterations and fall back to FP64 from the block reached so far.
    fp64_fallback = True
    eig_fast_maxit = 80
    sliver_ratio = 2e-3

    def __init__(self, vertices=None, tets=None, mode_num=16, mat=MatSet.Ceramic, order=1, mat_model=FixedLinear,
                 task=None, mesh_dir=None):
        if mesh_dir:
            self.mesh_dir = mesh_dir
            self.tetmesh = TetMesh.from_triangle_mesh(mesh_dir).to_high_order(order)
        else:
            if not vertices.is_cuda:
                raise RuntimeError("diffsound_b200: vertices/tets must be CUDA tensors (there is no CPU path)")
            self.tetmesh = TetMesh(vertices, tets).to_high_order(order)
        self.deform = Deform(self.tetmesh)
        if task == "mat_baseline":
            self.material_model = mat_model(Material(mat), baseline=True)
        else:
            self.material_model = mat_model(Material(mat))
        self.mode_num = mode_num
        self.U_hat_full = None
        self.task = task
        self._Kval = self._Mblk = None
        self._density_used = None
        self._lame_used = None
        self._X = None
        self._Xpad = None
        self._warm = []
        self._q = None
        self.eig_stats = None

    # ------------------------------------------------------------------ parameters
    def parameters(self):
        if self.task == "material":
            return self.material_model.parameters()
        elif self.task == "mat_baseline":
            return self.material_model.youngs.parameters()
        return None

    def init_material_coeffs(self, steps=5000, verbose=True):
        """Pre-train the (E, nu) logits towards the material table (diff_model.py:154-182)."""
        mm = self.material_model
        optimizer = torch.optim.Adam(mm.parameters(), lr=5e-3)
        gt_youngs, gt_poisson = mm.mat.youngs, mm.mat.poisson
        for _ in range(steps):
            optimizer.zero_grad()
            loss = (mm.youngs() - gt_youngs) ** 2 / gt_youngs ** 2 + (mm.poisson() - gt_poisson) ** 2 / gt_poisson ** 2
            loss.backward()
            optimizer.step()
        if verbose:
            print("(net) youngs: ", mm.youngs(), "poisson: ", mm.poisson())
            print("(material table) youngs: ", mm.mat.youngs, "poisson: ", mm.mat.poisson)
        self.scale = torch.eye(3, dtype=torch.float64, device=self.tetmesh.device)

    # ------------------------------------------------------------------ assembly
    def _assemble(self, density):
        d = self.deform
        mu, lam = self.material_model.lame()
        mu, lam = float(mu), float(lam)
        self._verts32 = d.verts_f32()
        pat = d.pattern
        if "_geom" not in self.__dict__:
            self._geom = torch.empty(d.num_tets * 14, dtype=torch.float64, device=d.device)
        self._Kval, self._Mblk = native.assemble_km(self._verts32, d.tets_i32, self.tetmesh.order, pat, mu, lam, d.ctab,
                                                    d.mtab(density), Kval=self._Kval, Mblk=self._Mblk, geom=self._geom)
        self._density_used = float(density)
        self._lame_used = (mu, lam)
        self.__dict__.pop("_stiff_coo", None)
        self.__dict__.pop("_mass_coo", None)

    def update_stiff_matrix(self, assemble_batch_size=20000):
        """K into the fixed pattern.  `assemble_batch_size` only bounded the reference's memory
        (diff_model.py:192-199); it is accepted and ignored."""
        dens = self._density_used if self._density_used is not None else self.material_model.mat.density
        self._assemble(dens)

    def update_mass_matrix(self, density):
        self._assemble(density)

    def _coo(self, values):
        idx = self.deform.pattern.coo_indices()
        n = self.deform.pattern.n
        return torch.sparse_coo_tensor(idx, values, (n, n), is_coalesced=True, check_invariants=False)

    @property
    def stiff_matrix(self):
        """torch sparse COO (n, n) fp64, coalesced, indices sorted by (row, col) -- what the reference
        keeps after `.coalesce()` (diff_model.py:216-220).  Materialised only when asked for."""
        if "_stiff_coo" not in self.__dict__:
            self._stiff_coo = self._coo(self._Kval)
        return self._stiff_coo

    @property
    def mass_matrix(self):
        if "_mass_coo" not in self.__dict__:
            self._mass_coo = self._coo(native.mass_expand(self.deform.pattern, self._Mblk))
        return self._mass_coo

    # ------------------------------------------------------------------ operators
    def stiff_func(self, x_in: torch.Tensor):
        """K(theta) x, differentiable in (E, nu) and in x: K = mu K_mu + lam K_lam (diff_model.py:314-328).
        Accepts (n,) or (n, k) like the reference, any number of columns."""
        x = x_in.unsqueeze(1) if x_in.dim() == 1 else x_in
        mu, lam = self.material_model.lame()
        mu = torch.as_tensor(mu, dtype=torch.float32).to(x.device)
        lam = torch.as_tensor(lam, dtype=torch.float32).to(x.device)
        force = _StiffFunc.apply(x, mu, lam, self)
        return force.squeeze(1) if x_in.dim() == 1 else force

    def _unit_stiffness(self):
        if "_Kunit" not in self.__dict__:
            d = self.deform
            v32 = d.verts_f32()
            mt = d.mtab(self.material_model.mat.density)
            Kmu, _ = native.assemble_km(v32, d.tets_i32, self.tetmesh.order, d.pattern, 1.0, 0.0, d.ctab, mt)
            Kla, _ = native.assemble_km(v32, d.tets_i32, self.tetmesh.order, d.pattern, 0.0, 1.0, d.ctab, mt)
            self._Kunit = (Kmu, Kla)
        return self._Kunit

    # ------------------------------------------------------------------ eigen-solve
    def eigen_decomposition(self):
        self._assemble(self.material_model.mat.density)
        self.eigen_decomposition_arpack()

    def _start_block(self, m, batch=0):
        """Start block of batch `batch` (n, m): the previous decomposition's Ritz block of this mesh when there is
        one (warm start; COPIED -- ds_lobpcg overwrites its block in place and earlier U_hat views / pending
        backward passes keep the old storage), else the six analytic rigid-body modes + seeded random columns."""
        n = self.deform.pattern.n
        dev = self.deform.device
        prev = self._warm[batch] if batch < len(self._warm) else None
        if prev is not None and prev.shape == (n, m):
            return prev.clone(), True
        g = torch.Generator(device=dev).manual_seed(batch)
        X = torch.randn(n, m, dtype=torch.float64, device=dev, generator=g)
        if batch == 0:
            p = self._verts32.double()
            p = p - p.mean(0, keepdim=True)
            X[:, :6] = 0
            for c in range(3):
                X[c::3, c] = 1
            X[0::3, 3], X[1::3, 3] = -p[:, 1], p[:, 0]
            X[1::3, 4], X[2::3, 4] = -p[:, 2], p[:, 1]
            X[2::3, 5], X[0::3, 5] = -p[:, 0], p[:, 2]
        return X, False

    def _has_slivers(self):
        """True when the flattest element of the mesh (smallest height over longest edge, from the corner nodes) is
        `sliver_ratio` times flatter than the median element."""
        if "_sliver" not in self.__dict__:
            with torch.no_grad():
                t = self.tetmesh.tets
                c = t[:, [0, 2, 4, 9]] if self.tetmesh.order == 2 else t
                p = self.tetmesh.vertices.detach()[c].double()
                e = p[:, [1, 2, 3, 2, 3, 3]] - p[:, [0, 0, 0, 1, 1, 2]]
                lmax = e.norm(dim=2).max(dim=1).values
                vol = torch.einsum("ij,ij->i", e[:, 0], torch.cross(e[:, 1], e[:, 2], dim=1)).abs() / 6
                q = vol / lmax ** 3
                self._sliver = bool(q.min() < self.sliver_ratio * q.median())
        return self._sliver

    #: pairs one eigensolver call has to converge at most (block 48 = 40 wanted + 8 guard columns); larger
    #: requests (geometry_train.py:147: mode_num = 64) are solved in batches with the converged vectors locked
    max_batch = 40

    @staticmethod
    def _block_for(need):
        return next((c for c in (16, 32, 48) if c >= need + min(4, c // 8)), None)

    def eigen_decomposition_arpack(self):
        """Lowest mode_num + 6 eigenpairs of K u = lambda M u, rigid six dropped.  The name is the
        reference's (diff_model.py:335-369: SciPy ARPACK shift-invert on the CPU); the solver is the
        device-resident LOBPCG of csrc/lobpcg.cu.  More pairs than one 48-column block holds are found in
        batches: every batch runs M-orthogonal to the pairs already converged (ds_lobpcg_opts.locked) and
        starts from the previous batch's guard columns."""
        k = self.mode_num
        need = k + 6
        pat = self.deform.pattern
        if pat.n < 3 * 16 or pat.n < need + 16:
            raise ValueError(f"mesh too small for {k} modes (n={pat.n})")
        deg = int(self.cheb_degree) or int(min(40, max(8, round(pat.n ** (1.0 / 3.0) / 3.0))))
        kw = {}
        coarse = self.deform.coarse if self.two_level else None
        if coarse is not None and 3 * coarse.n_nodes >= 3 * 48:
            # two-level p-multigrid preconditioner: P1 operator of the same mesh, same material
            mu, la = self._lame_used
            coarse.assemble(self._verts32, mu, la, coarse.ctab, self.deform.coarse_mtab(self._density_used))
            # n_c^(1/3) scaling fits compact bodies; thin shells (the bowl fixture) need the floor: 37 -> 28 outer iterations
            cdeg = int(self.coarse_degree) or int(min(64, max(32, round((3 * coarse.n_nodes) ** (1.0 / 3.0) / 1.2))))
            cratio = float(self.coarse_ratio) or 0.4 * cdeg * cdeg
            kw = dict(coarse=coarse, smooth_steps=int(self.smooth_steps), smooth_ratio=float(self.smooth_ratio),
                      coarse_degree=cdeg, coarse_ratio=cratio, nested_tol=self.nested_tol, nested_degree=self.nested_degree)
        found_X, found_lam, warm, all_stats = [], [], [], []
        done, batch, guard = 0, 0, None
        while done < need:
            want = min(need - done, self.max_batch)
            m = self._block_for(want)
            while pat.n < 3 * m and m > 16:
                m -= 16
            if m < want:
                raise ValueError(f"mesh too small for {k} modes (n={pat.n})")
            X, is_warm = self._start_block(m, batch)
            if guard is not None and not is_warm:
                gcols = min(guard.shape[1], m)
                X[:, :gcols] = guard[:, :gcols]            # approximate next modes left over from the previous batch
            locked = None
            if done:
                q = (done + 15) // 16 * 16
                locked = torch.zeros(pat.n, q, dtype=torch.float64, device=X.device)
                locked[:, :done] = torch.cat(found_X, dim=1)
            if kw:
                kw["nested"] = self.nested_start and not is_warm and batch == 0
            common = dict(nev=want, tol=self.eig_tol, n_rigid=6 if batch == 0 else 0,
                          coords=self._verts32 if self.morton else None, locked=locked)
            stats = None
            if not (self.fp64_fallback and self._has_slivers()):
                try:
                    lam, res, stats = native.lobpcg(pat, self._Kval, self._Mblk, X, cheb_degree=deg, cheb_ratio=0.4 * deg * deg,
                                                    maxit=min(self.eig_maxit, self.eig_fast_maxit) if self.fp64_fallback
                                                    else self.eig_maxit, **common, **kw)
                except RuntimeError:
                    if not self.fp64_fallback:
                        raise
                    stats = None
                if stats is not None and stats["status"] != 0 and self.fp64_fallback:
                    fast_stats, stats = stats, None
                    if not bool(torch.isfinite(X).all()):
                        X, _ = self._start_block(m, batch)
            if stats is None:
                deg64 = int(self.cheb_degree) or int(min(40, max(12, round(pat.n ** (1.0 / 3.0) / 1.5))))
                lam, res, stats = native.lobpcg(pat, self._Kval, self._Mblk, X, cheb_degree=deg64, cheb_ratio=0.4 * deg64 * deg64,
                                                maxit=self.eig_maxit, precond_fp64=True, **common)
                stats["precond"] = "fp64"
            if stats["status"] != 0:
                raise RuntimeError(f"eigensolver did not converge (batch {batch}): {stats}, "
                                   f"max residual {float(res[:want].max()):.3e}")
            all_stats.append(stats)
            warm.append(X)
            found_X.append(X[:, :want])
            found_lam.append(lam[:want])
            guard = X[:, want:]
            if batch == 0:
                self.ritz_values = lam          # all block columns of the first batch (rigid six, modes, guard columns)
            done += want
            batch += 1
        self._warm = warm
        self.eig_stats = all_stats[0] if len(all_stats) == 1 else dict(
            all_stats[0], batches=len(all_stats), iterations=sum(s["iterations"] for s in all_stats),
            spmm=sum(s["spmm"] for s in all_stats), per_batch=all_stats)
        if len(found_X) == 1:
            self._X = warm[0]                    # (n, m): wanted columns + guard columns
            self.U_hat_full = self._X[:, :need]
            lam_all = found_lam[0]
        else:
            lam_all, order = torch.sort(torch.cat(found_lam))
            self.U_hat_full = torch.cat(found_X, dim=1)[:, order].contiguous()
            self._X = self.U_hat_full
        self.eigenvalues = lam_all[6:need].clone()
        self.U_hat = self.U_hat_full[:, 6:need]
        # all wanted columns zero-padded to a multiple of 16 for the SpMM of get_vals()
        if self._X.shape[1] % 16 == 0:
            self._Xpad = self._X
        else:
            self._Xpad = torch.zeros(pat.n, (need + 15) // 16 * 16, dtype=torch.float64, device=self._X.device)
            self._Xpad[:, :need] = self.U_hat_full
        self._q = None

    # ------------------------------------------------------------------ differentiable outputs
    def _material_forms(self):
        if self._q is None:
            d = self.deform
            self._q = native.eigval_quadforms_material(self._verts32, d.tets_i32, self.tetmesh.order,
                                                       d.mtab(self._density_used), d.wsum, self.U_hat)
        return self._q

    def get_undamped_freqs(self):
        """(mode_num, 1) fp32 undamped frequencies with gradient to the (E, nu) logits
        (diff_model.py:371-388)."""
        dev = self.deform.device
        predict = torch.zeros(self.mode_num, dtype=torch.float32, device=dev)
        predict += self.eigenvalues
        if self.task != "gt":
            q_mu, q_la, q_m = self._material_forms()
            mu, lam = self.material_model.lame()
            mu, lam = torch.as_tensor(mu).to(dev), torch.as_tensor(lam).to(dev)
            # fp64 until the final cast (the reference does this sum in fp32, diff_model.py:382-386)
            add_term = (mu.double() * q_mu + lam.double() * q_la) - self.eigenvalues * q_m
            predict = predict + add_term.float()
        return (torch.sqrt(predict) / 2 / np.pi).unsqueeze(1)

    def get_vals(self):
        """(mode_num, 1) fp32 eigenvalues, differentiable w.r.t. the vertex positions
        (diff_model.py:390-399)."""
        return _EigvalShape.apply(self.tetmesh.vertices, self)
