"""Operator-form LOBPCG worker: A (and B, iK) given as callables, dense or arbitrary sparse tensors.

API mirror of the reference's worker class `LOBPCG` (src/lobpcg/_lobpcg.py:214-686): same constructor
arguments, the same public state (`A, B, iK, iparams, fparams, bparams, X, E, R, S, tvars, ivars, fvars,
bvars`), the same `run / update / update_residual / update_converged_count / stop_iteration /
call_tracker` methods, the same convergence test (`_lobpcg.py:307-333`) and the same hooks
(`tracker(worker)` after every step, `worker.bvars['force_stop']`, `profiler` = a tensorboard trace
directory, `_lobpcg.py:344-376`).

This is the path `lobpcg_func` exists for (`_lobpcg.py:123-212`: "A doesn't need to be a matrix"): the
operator products are whatever the caller hands in; everything else of a step -- the fused Gram pair
S^T(AS), S^T(BS), the small generalised eigen-solve, the basis update -- runs on the library's FP64
kernels (`ds_gram_sym2_f64`, `ds_eigh_generalized_f64`, `ds_block_gemm_f64`).  K/M with the FEM block
structure do not come here: `lobpcg()` routes them to the device-resident driver `ds_lobpcg`.

Both of the reference's methods ('basic', 'ortho') map to one update: Rayleigh-Ritz on [X, W, P] with W
B-orthogonalised against X; a failed Cholesky of the projected B drops P for that step.
"""
import torch

from .. import native


def _is_op(a):
    return callable(a) and not torch.is_tensor(a)


class LOBPCG(object):
    def __init__(self, A, B, X, E, iK, iparams, fparams, bparams, method, tracker, profiler):
        self.A, self.B, self.iK = A, B, iK
        self.iparams, self.fparams, self.bparams = iparams, fparams, bparams
        self.method, self.tracker, self.profiler = method, tracker, profiler
        m, n = iparams["m"], iparams["n"]
        self.X = X
        self.E = E if E is not None else torch.zeros((n,), dtype=X.dtype, device=X.device)
        self.R = torch.zeros((m, n), dtype=X.dtype, device=X.device)
        self.S = torch.zeros((m, 3 * n), dtype=X.dtype, device=X.device)
        self.tvars, self.ivars, self.fvars, self.bvars = {}, {"istep": 0}, {"_": 0.0}, {"_": False}
        # fp64 work blocks, columns [X | W | P], width padded to a multiple of 8 (DMMA tiles)
        self._w = (n + 7) // 8 * 8
        if 3 * self._w > 144:
            raise NotImplementedError(f"lobpcg: block size {n} exceeds 48 columns (Rayleigh-Ritz limit 144)")
        dev = X.device
        f64 = dict(dtype=torch.float64, device=dev)
        ld = 3 * self._w
        self._S, self._AS, self._BS = (torch.zeros(m, ld, **f64) for _ in range(3))
        self._S2, self._AS2, self._BS2 = (torch.zeros(m, ld, **f64) for _ in range(3))
        self._GK = torch.zeros(ld, ld, **f64)
        self._GM = torch.zeros(ld, ld, **f64)
        self._np = 0
        self._sign = -1.0 if bparams.get("largest", True) else 1.0

    # ------------------------------------------------------------------ operators
    def _apply(self, op, V):
        """op V in fp64; V is fp64 (m, c).  Tensors are applied in their own dtype, callables in X's."""
        if op is None:
            return V
        if _is_op(op):
            return op(V.to(self.X.dtype)).to(torch.float64)
        Vc = V.to(op.dtype)
        out = torch.sparse.mm(op, Vc) if op.layout != torch.strided else op @ Vc
        return out.to(torch.float64)

    def _A(self, V):
        return self._sign * self._apply(self.A, V)      # largest=True: lowest pairs of -A

    # ------------------------------------------------------------------ reference-shaped pieces
    def update_residual(self):
        n = self.iparams["n"]
        AX = self._sign * self._AS[:, :n]
        self.R = (AX - self._BS[:, :n] * self.E.to(torch.float64)).to(self.X.dtype)

    def update_converged_count(self):
        prev = self.ivars["converged_count"]
        tol = self.fparams["tol"]
        A_norm, B_norm = self.fvars["A_norm"], self.fvars["B_norm"]
        E, X, R = self.E, self.X, self.R
        rerr = torch.norm(R, 2, (0,)) * (torch.norm(X, 2, (0,)) * (A_norm + E[:X.shape[-1]].abs() * B_norm)) ** -1
        count = 0
        for b in (rerr < tol).tolist():
            if not b:
                break
            count += 1
        count = max(count, prev)          # soft: a pair that drifted back above tol keeps its place
        self.ivars["converged_count"] = count
        self.tvars["rerr"] = rerr
        return count

    def stop_iteration(self):
        return (self.bvars.get("force_stop", False) or self.ivars["iterations_left"] == 0
                or self.ivars["converged_count"] >= self.iparams["k"])

    def call_tracker(self):
        if self.tracker is not None:
            self.tracker(self)

    def run(self):
        self.call_tracker()
        self.update()
        self.call_tracker()
        if self.profiler:
            with torch.profiler.profile(
                    schedule=torch.profiler.schedule(wait=1, warmup=1, active=3, repeat=1),
                    on_trace_ready=torch.profiler.tensorboard_trace_handler(self.profiler),
                    record_shapes=True, profile_memory=True, with_stack=True) as prof:
                while not self.stop_iteration():
                    self.update()
                    prof.step()
                    self.call_tracker()
        else:
            while not self.stop_iteration():
                self.update()
                self.call_tracker()

    # ------------------------------------------------------------------ one step
    def _rayleigh_ritz(self, ncols, idx=None):
        """Lowest pairs of the pencil projected on the work columns `idx` (default: the first `ncols`).
        Returns (theta, C) with C's rows at the work-column positions (zero rows elsewhere), or None."""
        w8 = ncols // 8
        native.gram_sym2(self._S, self._AS, self._BS, list(range(w8)), self._GK, self._GM)
        GK, GM = self._GK[:ncols, :ncols], self._GM[:ncols, :ncols]
        GK = torch.triu(GK) + torch.triu(GK, 1).T                 # only the upper-triangle tiles are written by the kernel
        GM = torch.triu(GM) + torch.triu(GM, 1).T
        if idx is not None:
            GK, GM = GK[idx][:, idx].contiguous(), GM[idx][:, idx].contiguous()
        # shift that makes the scaled projected A positive definite: Gershgorin bound of D GK D, D = diag(GM)^-1/2
        d = torch.rsqrt(torch.clamp(torch.diagonal(GM), min=1e-300))
        sig = float((GK * d[:, None] * d[None, :]).abs().sum(1).max()) * 1.0000001 + 1e-300
        for _ in range(6):
            theta, C, info = native.eigh_generalized(GK, GM, sig)
            code = int(info[0])
            if code == 0:
                if idx is not None:
                    Cf = torch.zeros(ncols, C.shape[1], dtype=C.dtype, device=C.device)
                    Cf[idx] = C
                    C = Cf
                return theta, C
            if code < 1000:              # projected B is not positive definite: dependent search directions
                return None
            sig *= 8.0                   # projected A + sig B not yet positive definite (indefinite A)
        return None

    def update(self):
        n, w = self.iparams["n"], self._w
        f64 = torch.float64
        S, AS, BS = self._S, self._AS, self._BS
        if self.ivars["istep"] == 0:
            X0 = torch.randn_like(self.X)
            iX = float(torch.norm(X0)) ** -1
            self.fvars["X_norm"] = 1.0 / iX
            self.fvars["A_norm"] = float(torch.norm(self._apply(self.A, X0.to(f64)))) * iX
            self.fvars["B_norm"] = float(torch.norm(self._apply(self.B, X0.to(f64)))) * iX
            self.ivars["iterations_left"] = self.iparams["niter"]
            self.ivars["converged_count"] = 0
            self.ivars["converged_end"] = 0
            self._best = None
            S.zero_(); AS.zero_(); BS.zero_()
            S[:, :n] = self.X.to(f64)
            if w > n:                      # padding columns: extra search directions
                g = torch.Generator(device=S.device).manual_seed(0x5EED0B200)
                S[:, n:w] = torch.randn(S.shape[0], w - n, dtype=f64, device=S.device, generator=g)
            AS[:, :w] = self._A(S[:, :w])
            BS[:, :w] = self._apply(self.B, S[:, :w])
            rr = self._rayleigh_ritz(w)
            if rr is None:
                raise ValueError("lobpcg: the initial block X is not B-independent")
            ncols = w
        else:
            # W = iK R for the columns that have not converged (the leading `converged_count` pairs are locked softly:
            # they stay in the Rayleigh-Ritz basis but get no new search directions, _lobpcg.py:394-431), B-orthogonalised
            # against X
            nc = self.ivars["converged_count"]
            Rw = torch.zeros(S.shape[0], w, dtype=f64, device=S.device)
            Rw[:, nc:n] = self.R.to(f64)[:, nc:]
            if w > n:
                Rw[:, n:] = AS[:, n:w] - BS[:, n:w] * self._theta[n:w]
            W = self._apply(self.iK, Rw).contiguous()
            coef = native.gram(BS[:, :w], W)
            native.block_gemm(S[:, :w], -coef, beta=1.0, out=W)
            S[:, w:2 * w] = W
            AS[:, w:2 * w] = self._A(W)
            BS[:, w:2 * w] = self._apply(self.B, W)
            act = torch.arange(nc, w, device=S.device)
            rr = None
            if self._np:
                ncols = 3 * w
                rr = self._rayleigh_ritz(ncols, torch.cat([torch.arange(w, device=S.device), w + act, 2 * w + act]))
            if rr is None:
                ncols = 2 * w
                rr = self._rayleigh_ritz(ncols, torch.cat([torch.arange(w, device=S.device), w + act]))
            if rr is None:                 # search directions collapsed: nothing more to gain
                self.ivars["iterations_left"] = 1
                ncols, rr = w, self._rayleigh_ritz(w)
                if rr is None:
                    raise RuntimeError("lobpcg: Rayleigh-Ritz breakdown")
        theta, C = rr
        S2, AS2, BS2 = self._S2, self._AS2, self._BS2
        C1 = C[:, :w].contiguous()
        for src, dst in ((S, S2), (AS, AS2), (BS, BS2)):
            dst.zero_()
            native.block_gemm(src[:, :ncols], C1, out=dst[:, :w])
            if ncols > w:                  # P' = [W P] C[w:, :w]
                native.block_gemm(src[:, w:ncols], C[w:ncols, :w].contiguous(), out=dst[:, 2 * w:3 * w])
        self._np = w if ncols > w else 0
        self._S, self._AS, self._BS, self._S2, self._AS2, self._BS2 = S2, AS2, BS2, S, AS, BS
        self._theta = theta[:w].clone()
        self.E = (self._sign * theta[:n]).to(self.X.dtype)
        self.X = self._S[:, :n].to(self.X.dtype)
        self.update_residual()
        self.update_converged_count()
        self.S[:, :n] = self.X
        self.ivars["iterations_left"] -= 1
        self.ivars["istep"] += 1
        # Divergence guard: in low precision (fp32 operators) the attainable residual can sit above `tol`; iterating on
        # noise-level residuals then degrades the basis.  Keep the best state seen (largest converged count, then
        # smallest leading residual) and stop with it once the leading residual has grown tenfold over its best.
        k = self.iparams["k"]
        score = (self.ivars["converged_count"], -float(self.tvars["rerr"][:k].max()))
        if self._best is None or score >= self._best[0]:
            self._best = (score, self.E.clone(), self.X.clone(), self.tvars["rerr"].clone(), self.ivars["converged_count"])
        elif -score[1] > 10.0 * -self._best[0][1] or self.ivars["iterations_left"] == 0:
            _, self.E, self.X, rerr, cc = self._best
            self.tvars["rerr"] = rerr
            self.ivars["converged_count"] = cc
            self.bvars["force_stop"] = True
