"""Round-2 planning probe (not a bench): how long do the library eigensolvers and an fp64 GEMM of the K8 size take?"""
import json, time, torch
torch.manual_seed(0)
d, n = 2000, 20000
Z = torch.randn(n, d, device="cuda", dtype=torch.float64)
C = (Z.T @ Z) / n
out = {}
def t(f, reps=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
out["eigh_f64_ms"] = t(lambda: torch.linalg.eigh(C))
C32 = C.float()
out["eigh_f32_ms"] = t(lambda: torch.linalg.eigh(C32))
out["eigvalsh_f64_ms"] = t(lambda: torch.linalg.eigvalsh(C))
out["gemm_f64_2000_ms"] = t(lambda: C @ C, 10)
out["gemm_f32_2000_ms"] = t(lambda: C32 @ C32, 10)
out["cholesky_f64_ms"] = t(lambda: torch.linalg.cholesky(C), 5)
out["qr_f64_2000x128_ms"] = t(lambda: torch.linalg.qr(C[:, :128]), 5)
# accuracy of fp32 eigenvectors + one Ogita-Aishima refinement step in fp64 (top 50)
w, V = torch.linalg.eigh(C)
w32, V32 = torch.linalg.eigh(C32)
X = V32.double()
def refine(X):
    R = torch.eye(d, device="cuda", dtype=torch.float64) - X.T @ X
    S = X.T @ C @ X
    lam = torch.diagonal(S) / (1 - torch.diagonal(R))
    D = lam[None, :] - lam[:, None]
    E = (S + R * lam[None, :]) / D
    E.fill_diagonal_(0)
    E = E + torch.diag(torch.diagonal(R) / 2)
    return X + X @ E, lam
def err(X):
    Xt, Vt = X[:, -50:], V[:, -50:]
    s = torch.sign((Xt * Vt).sum(0))
    return float((Xt * s - Vt).abs().max())
out["top50_vec_err_f32"] = err(X)
X1, lam1 = refine(X)
out["top50_vec_err_refined1"] = err(X1)
X2, lam2 = refine(X1)
out["top50_vec_err_refined2"] = err(X2)
out["refine_step_ms"] = t(lambda: refine(X), 3)
out["min_rel_gap_top50"] = float(((w[-50:] - w[-51:-1]) / w[-1]).min())
print(json.dumps(out))
