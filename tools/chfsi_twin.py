"""NumPy twin of csrc/eig.cu: chfsi_topk — the same flow, constants and update rules, used to validate the algorithm on
spectra a GPU box is not needed for (tests/test_chfsi_twin.py) and to explore tunables before spending GPU time.

    bounds (L = 16 Krylov steps, CGS2) -> [filter degree m (amplification of the top capped at 1e8, trimmed to the
    degree that reaches the target in R rounds) + CholeskyQR2] x R
    -> Rayleigh-Ritz -> residuals of the top k -> new cut / bounds -> ...            (at most 6 outer rounds)
"""
from __future__ import annotations

import numpy as np

AMP_CAP, MAX_DEGREE, MAX_ROUNDS, MAX_OUTER, MAX_PRODUCTS = 1e8, 32, 24, 6, 400


def cheb_filter(mul, Y, m, lo, cut, up):
    """Scaled Chebyshev recurrence (Zhou & Saad): bounded by 1 on [lo, cut], growing above; `up` only scales."""
    e, c = (cut - lo) / 2, (cut + lo) / 2
    sigma1 = e / (up - c)
    sigma = sigma1
    Yp, Yc = Y, (mul(Y) - c * Y) * (sigma1 / e)
    for _ in range(2, m + 1):
        sn = 1.0 / (2.0 / sigma1 - sigma)
        Yp, Yc, sigma = Yc, (mul(Yc) - c * Yc) * (2 * sn / e) - (sigma * sn) * Yp, sn
    return Yc


def krylov_bounds(C, rng, L):
    d = C.shape[0]
    V = np.zeros((d, L + 1))
    v = rng.uniform(-1, 1, d)
    V[:, 0] = v / np.linalg.norm(v)
    beta, ref = 0.0, None
    for j in range(L):
        w = C @ V[:, j]
        if j == 0:
            ref = np.linalg.norm(w)
        for _ in range(2):
            w -= V[:, :j + 1] @ (V[:, :j + 1].T @ w)
        beta = np.linalg.norm(w)
        if not (beta > 1e-13 * abs(ref)) or not np.isfinite(beta):
            return None  # Krylov breakdown (scale_by_inv_norm_kernel raises the flag): the caller falls back
        V[:, j + 1] = w / beta
    H = V[:, :L].T @ (C @ V[:, :L])
    th, S = np.linalg.eigh((H + H.T) / 2)
    return th, S, beta


def dos_quantile(th, w, up, p):
    """Eigenvalue below which a fraction 1 - p of the spectrum lies, from Ritz values th (ascending) with weights w: CDF from
    the top by the midpoint rule, linear interpolation between the nodes (eig.cu: dos_quantile)."""
    t, ww = th[::-1], w[::-1]
    c = np.cumsum(ww) - ww / 2
    if p <= c[0]:
        return up - (up - t[0]) * p / c[0]
    for i in range(len(t) - 1):
        if c[i] <= p < c[i + 1]:
            return t[i] + (p - c[i]) / (c[i + 1] - c[i]) * (t[i + 1] - t[i])
    return t[-1]


def block_width(d, k):
    return min(d // 4, ((max(3 * k, k + 96) + 31) // 32) * 32)


def chfsi_topk(C, k, seed=0, L=16, target=1e11, tol=1e-11, b=None, log=None):
    """Returns (eigenvalues ascending, eigenvectors, stats) or None where the CUDA code would fall back to syevd."""
    rng = np.random.default_rng(seed)
    d = C.shape[0]
    b = b or block_width(d, k)
    if b < k + 16 or d < 512:
        return None
    kb = krylov_bounds(C, rng, L)
    if kb is None:
        return None
    th, S, beta = kb
    if not np.all(np.isfinite(th)):
        return None
    span = th[-1] - th[0]
    if not span > 0:
        return None
    lo = th[0] - beta * abs(S[L - 1, 0]) - 0.01 * span
    up = th[-1] + beta * abs(S[L - 1, L - 1])
    cut, lamk, cw, have_k, have_cut = th[0], th[-1], 0.0, False, False
    for i in range(L - 1, -1, -1):  # density of states from the Ritz weights, walking down from the top
        cw += S[0, i] ** 2
        if not have_k and cw >= k / d:
            lamk, have_k = th[i], True
        if not have_cut and cw >= 0.8 * b / d:
            cut, have_cut = th[i], True
    # the step function above can cross k / d one Ritz value too early (an optimistic lamk makes the first outer round
    # under-filter); the midpoint-rule, linearly interpolated quantile is smoother: take the lower of the two
    lamk = min(lamk, dos_quantile(th, S[0, :] ** 2, up, k / d))
    cut = max(min(cut, th[-1] - 0.02 * span), lo + 0.05 * span)
    lamk = max(lamk, cut + 0.01 * span)
    Y = rng.uniform(-1, 1, (d, b))
    stats = dict(block_products=0, cholqr=0, outer=0)
    for outer in range(MAX_OUTER):
        stats["outer"] += 1
        e, c = (cut - lo) / 2, (cut + lo) / 2
        if not (e > 0 and np.isfinite(e)):
            return None
        xtop, xk = max((up - c) / e, 1 + 1e-12), max((lamk - c) / e, 1 + 1e-9)
        m = int(max(2, min(MAX_DEGREE, np.floor(np.arccosh(AMP_CAP) / np.arccosh(xtop)))))
        amp = np.cosh(m * np.arccosh(xk))
        # later outer rounds only top up what the last Rayleigh-Ritz step showed missing (x10 margin), eig.cu: target_now
        tgt = target if outer == 0 else float(min(target, max(1e3, 1e3 * stats["max_residual"] / tol)))
        R = int(np.ceil(np.log(tgt) / np.log(max(amp, 1.0001))))
        R = max(1, min(R, 3 if outer == 0 else MAX_ROUNDS))
        # the smallest degree that reaches the target in exactly R rounds (eig.cu: m_use)
        m = max(2, min(m, int(np.ceil(np.arccosh(tgt ** (1.0 / R)) / np.arccosh(xk)))))
        if stats["block_products"] + R * m > MAX_PRODUCTS:
            return None  # work budget: costlier than the syevd it replaces
        for _r in range(R):
            Y = cheb_filter(lambda X: C @ X, Y, m, lo, cut, up)
            stats["block_products"] += m
            try:
                # one CholeskyQR pass between rounds while the round amplified the top by <= 1e7, two before Rayleigh-Ritz
                for _ in range(2 if (_r + 1 == R or np.cosh(m * np.arccosh(xtop)) > 1e7) else 1):
                    Rm = np.linalg.cholesky(Y.T @ Y).T
                    Y = np.linalg.solve(Rm.T, Y.T).T
            except np.linalg.LinAlgError:
                return None
            stats["cholqr"] += 1
        W = C @ Y
        stats["block_products"] += 1
        G = Y.T @ W
        tt, Sg = np.linalg.eigh((G + G.T) / 2)
        Y2, CW = Y @ Sg, W @ Sg[:, b - k:]
        res = np.linalg.norm(CW - Y2[:, b - k:] * tt[b - k:], axis=0)
        if not np.all(np.isfinite(res)):
            return None
        rel = res.max() / max(abs(tt[-1]), 1e-300)
        if log is not None:
            log.append(dict(outer=outer, lo=lo, cut=cut, up=up, lamk=lamk, degree=m, rounds=R, amp=amp, residual=rel))
        Y = Y2
        stats["max_residual"] = rel
        if rel <= tol:
            return tt[b - k:], Y[:, b - k:], stats
        # the block's smallest Ritz value raises the cut when the first one left more than b eigenvalues above it; it only
        # lowers it when the k-th Ritz value shows the cut sat above wanted eigenvalues (a badly converged last vector has a
        # Ritz value deep inside the bulk, which would blunt the next filter)
        lamk, up = tt[b - k], max(up, tt[-1])
        cut = tt[0] if lamk <= cut else max(cut, tt[0])
        if not cut > lo:
            lo = cut - 0.05 * (up - cut)
    return None


def spectrum(kind, d, rng):
    """Symmetric test matrices: 'flat' (the bench: Marchenko-Pastur bulk only), 'spiked' (strong components over a bulk),
    'lowrank' (n < d Gram), 'clustered' (the k-th eigenvalue sits in a tight cluster), 'powerlaw'."""
    if kind in ("flat", "spiked"):
        n = 500.0 * d
        G = rng.standard_normal((d, d))
        C = n * (np.eye(d) + np.sqrt(d / n) * (G + G.T) / np.sqrt(2 * d))
        if kind == "spiked":
            U, _ = np.linalg.qr(rng.standard_normal((d, 12)))
            C = C + (U * (n * np.linspace(20, 0.5, 12))) @ U.T
        return C
    if kind == "lowrank":
        Z = rng.standard_normal((d // 3, d))
        return Z.T @ Z
    Q, _ = np.linalg.qr(rng.standard_normal((d, d)))
    if kind == "clustered":
        lam = np.concatenate([np.full(8, 10.0), 5 + 1e-6 * np.arange(12), np.linspace(1, 4.9, d - 20)])
    elif kind == "powerlaw":
        lam = 300.0 * np.arange(1, d + 1) ** -0.8 + 0.2
    else:
        raise ValueError(kind)
    return (Q * lam) @ Q.T


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for kind in ("flat", "spiked", "lowrank", "clustered", "powerlaw"):
        C = spectrum(kind, 2000, rng)
        w, V = np.linalg.eigh(C)
        for L, target in ((40, 1e11), (16, 1e10)):
            out = chfsi_topk(C, 50, L=L, target=target)
            if out is None:
                print(f"{kind:10s} L={L}: falls back to syevd")
                continue
            lam, X, st = out
            s = np.sign(np.sum(X * V[:, -50:], axis=0))
            print(f"{kind:10s} L={L} target={target:.0e}: {st}  eigenvalue err {np.abs(lam - w[-50:]).max() / abs(w[-1]):.1e}  "
                  f"vector err {np.abs(X * s - V[:, -50:]).max():.1e}")
