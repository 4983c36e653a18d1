import sys, torch, numpy as np
sys.path.insert(0, '.')
from diffsound_b200 import native
torch.manual_seed(0)
n = 400
for w in (8, 16, 48):
    ld = 3 * w
    S = torch.randn(n, ld, dtype=torch.float64, device='cuda'); KS = torch.randn_like(S); MS = S.clone()
    GK = torch.zeros(ld, ld, dtype=torch.float64, device='cuda'); GM = torch.zeros_like(GK)
    native.gram_sym2(S, KS, MS, list(range(w // 8)), GK, GM)
    rk = S[:, :w].T @ KS[:, :w]; rm = S[:, :w].T @ MS[:, :w]
    print('w', w, 'gram err', float((torch.triu(GK[:w, :w]) - torch.triu(rk)).abs().max()), float((torch.triu(GM[:w, :w]) - torch.triu(rm)).abs().max()))
    A = torch.randn(w, w, dtype=torch.float64, device='cuda'); A = A @ A.T + torch.eye(w, dtype=torch.float64, device='cuda')
    GKf = torch.zeros(ld, ld, dtype=torch.float64, device='cuda'); GKf[:w, :w] = A
    theta, C, info = native.eigh_generalized(GKf[:w, :w], GM[:w, :w], 10.0)
    import scipy.linalg as sla
    M = torch.triu(GM[:w, :w]); M = (M + torch.triu(M, 1).T).cpu().numpy()
    ref = sla.eigh(A.cpu().numpy(), M, eigvals_only=True)
    print('  eigh info', info.tolist(), 'err', np.abs(theta.cpu().numpy() - ref).max() / abs(ref).max())
