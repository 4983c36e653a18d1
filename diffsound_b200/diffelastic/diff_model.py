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


class _EigvalShape(torch.autograd.Function):
    """get_vals(): value lambda + (u^T K u - lambda u^T M u), gradient u^T (dK - lambda dM) u w.r.t.
    the (promoted) vertex positions."""

    @staticmethod
    def forward(ctx, vertices, obj):
        ctx.obj = obj
        X, lam = obj._X, obj.eigenvalues          # the whole eigensolver block (16 | 32 | 48 columns)
        pat = obj.deform.pattern
        KX, MX = native.spmm_k_and_m(pat, obj._Kval, obj._Mblk, X)
        lo, hi = 6, 6 + obj.mode_num
        uku = torch.diagonal(native.gram(X, KX))[lo:hi]
        umu = torch.diagonal(native.gram(X, MX))[lo:hi]
        predict = torch.zeros(obj.mode_num, dtype=torch.float32, device=X.device)
        predict += lam
        predict += uku - lam * umu
        return predict.unsqueeze(1)

    @staticmethod
    def backward(ctx, g):
        obj = ctx.obj
        d = obj.deform
        mu, lam_l = obj._lame_used
        inc_ptr, inc = d.incidence
        gv = g.reshape(-1).to(torch.float64).contiguous()
        grad = native.eigval_grad_shape(obj._verts32, d.tets_i32, obj.tetmesh.order, mu, lam_l, d.ctab,
                                        d.mtab(obj._density_used), obj.U_hat, obj.eigenvalues, gv, inc_ptr, inc)
        return grad.to(obj.tetmesh.vertices.dtype), None


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
    morton = True           # the FP32 preconditioner keeps its operator in a Morton node numbering (SpMM locality)
    eig_maxit = 400

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
        """K(theta) x, differentiable in (E, nu): K = mu K_mu + lam K_lam (diff_model.py:314-328)."""
        x = x_in.unsqueeze(1) if x_in.dim() == 1 else x_in
        Kmu, Kla = self._unit_stiffness()
        cols = x.shape[1]
        cp = (cols + 15) // 16 * 16
        xp = torch.zeros(x.shape[0], cp, dtype=torch.float64, device=x.device)
        xp[:, :cols] = x.detach()
        pat = self.deform.pattern
        ymu = native.spmm(pat, Kmu, None, xp)[:, :cols]
        yla = native.spmm(pat, Kla, None, xp)[:, :cols]
        mu, lam = self.material_model.lame()
        if torch.is_tensor(mu):
            mu, lam = mu.to(x.device), lam.to(x.device)
        force = (mu * ymu + lam * yla).to(x_in.dtype)
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

    def _start_block(self, m):
        n = self.deform.pattern.n
        dev = self.deform.device
        if self._X is not None and self._X.shape == (n, m):
            return self._X              # warm start from the previous decomposition of this mesh
        g = torch.Generator(device=dev).manual_seed(0)
        X = torch.randn(n, m, dtype=torch.float64, device=dev, generator=g)
        p = self._verts32.double()
        p = p - p.mean(0, keepdim=True)
        X[:, :6] = 0
        for c in range(3):
            X[c::3, c] = 1
        X[0::3, 3], X[1::3, 3] = -p[:, 1], p[:, 0]
        X[1::3, 4], X[2::3, 4] = -p[:, 2], p[:, 1]
        X[2::3, 5], X[0::3, 5] = -p[:, 0], p[:, 2]
        return X

    def eigen_decomposition_arpack(self):
        """Lowest mode_num + 6 eigenpairs of K u = lambda M u, rigid six dropped.  The name is the
        reference's (diff_model.py:335-369: SciPy ARPACK shift-invert on the CPU); the solver is the
        device-resident LOBPCG of csrc/lobpcg.cu."""
        k = self.mode_num
        need = k + 6
        m = next((c for c in (16, 32, 48) if c >= need + min(4, c // 8)), None)
        if m is None:
            raise NotImplementedError(f"mode_num={k}: the eigensolver block is limited to 48 columns (mode_num <= 42)")
        pat = self.deform.pattern
        if pat.n < 3 * m:
            raise ValueError(f"mesh too small for {k} modes (n={pat.n})")
        X = self._start_block(m)
        deg = int(min(40, max(8, round(pat.n ** (1.0 / 3.0) / 3.0))))
        kw = {}
        coarse = self.deform.coarse if self.two_level else None
        if coarse is not None and 3 * coarse.n_nodes >= 3 * m:
            # two-level p-multigrid preconditioner: P1 operator of the same mesh, same material
            mu, la = self._lame_used
            coarse.assemble(self._verts32, mu, la, coarse.ctab, self.deform.coarse_mtab(self._density_used))
            # n_c^(1/3) scaling fits compact bodies; thin shells (the bowl fixture) need the floor: 37 -> 28 outer iterations
            cdeg = int(self.coarse_degree) or int(min(64, max(32, round((3 * coarse.n_nodes) ** (1.0 / 3.0) / 1.2))))
            cratio = float(self.coarse_ratio) or 0.4 * cdeg * cdeg
            kw = dict(coarse=coarse, smooth_steps=int(self.smooth_steps), smooth_ratio=float(self.smooth_ratio),
                      coarse_degree=cdeg, coarse_ratio=cratio, nested=self.nested_start and self._X is not X,
                      nested_tol=self.nested_tol, nested_degree=self.nested_degree)
        lam, res, stats = native.lobpcg(pat, self._Kval, self._Mblk, X, nev=need, tol=self.eig_tol, maxit=self.eig_maxit,
                                        cheb_degree=deg, cheb_ratio=0.4 * deg * deg, n_rigid=6,
                                        coords=self._verts32 if self.morton else None, **kw)
        if stats["status"] != 0:
            raise RuntimeError(f"eigensolver did not converge: {stats}, max residual {float(res[:need].max()):.3e}")
        self.eig_stats = stats
        self._X = X
        self.U_hat_full = X[:, :need]
        self.eigenvalues = lam[6:need].clone()
        self.ritz_values = lam          # all block columns (rigid six, wanted modes, guard columns)
        self.U_hat = X[:, 6:need]
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
