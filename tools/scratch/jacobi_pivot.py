# scratch: does diagonal-pivoted Cholesky (Veselic-Hari / Drmac) cut the one-sided Jacobi sweep count on the pooled Grams?
import numpy as np, sys, torch
sys.path.insert(0, "/root/repo")
from oracle import synth
import dataclasses

def jacobi_sweeps(A, tol=3e-7, small=3e-5, max_sweeps=40, order="oddeven"):
    A = A.astype(np.float32).copy()
    n = A.shape[1]
    pos = list(range(n))
    for sweep in range(max_sweeps):
        big = False
        # odd-even transposition with swap: n steps
        for step in range(n):
            start = step % 2
            for i in range(start, n - 1, 2):
                p, q = pos[i], pos[i + 1]
                x, y = A[:, p], A[:, q]
                al, be, ga = x @ x, y @ y, x @ y
                if abs(ga) > tol * np.sqrt(al * be):
                    if abs(ga) > small * np.sqrt(al * be): big = True
                    d = be - al; h = 2 * ga
                    r = np.hypot(d, h)
                    c2 = 0.5 + 0.5 * abs(d) / r
                    cs = np.sqrt(c2); sn = 0.5 * h / r / cs
                    if d < 0: sn = -sn
                    A[:, p], A[:, q] = cs * x - sn * y, sn * x + cs * y
                pos[i], pos[i + 1] = pos[i + 1], pos[i]
        if not big:
            return sweep + 1
    return max_sweeps

def pivoted_chol(G):
    G = G.astype(np.float64).copy(); n = G.shape[0]
    perm = np.arange(n); L = np.zeros_like(G)
    d = np.diag(G).copy()
    for j in range(n):
        k = j + np.argmax(d[j:])
        if k != j:
            G[[j, k]] = G[[k, j]]; G[:, [j, k]] = G[:, [k, j]]
            L[[j, k]] = L[[k, j]]; d[[j, k]] = d[[k, j]]; perm[[j, k]] = perm[[k, j]]
        L[j, j] = np.sqrt(d[j])
        L[j + 1:, j] = (G[j + 1:, j] - L[j + 1:, :j] @ L[j, :j]) / L[j, j]
        d[j + 1:] -= L[j + 1:, j] ** 2
    return L, perm

w = dataclasses.replace(synth.CONFIGS["cfg2"], B=64)
inp = synth.make_inputs(w)
torch.manual_seed(0)
pt = torch.empty(w.Ds, w.Dt); torch.nn.init.orthogonal_(pt)
for j in (0, 5, 11):
    Z = (inp["teacher"][j].float().reshape(-1, w.Dt) @ pt.T).double().numpy()
    Zc = Z - Z.mean(0)
    G = Zc.T @ Zc
    L = np.linalg.cholesky(G)
    Lp, perm = pivoted_chol(G)
    print(f"teacher layer {j}: sweeps on G {jacobi_sweeps(G)}, on chol(G) {jacobi_sweeps(L)}, on pivoted chol {jacobi_sweeps(Lp)}, on L^T (rows) {jacobi_sweeps(L.T.copy())}, pivoted L^T {jacobi_sweeps(Lp.T.copy())}", flush=True)
S = inp["student"][sorted(inp["student"])[0]].float().reshape(-1, w.Ds).double().numpy()
Sc = S - S.mean(0); G = Sc.T @ Sc
L = np.linalg.cholesky(G); Lp, perm = pivoted_chol(G)
print(f"student: sweeps on G {jacobi_sweeps(G)}, chol {jacobi_sweeps(L)}, pivoted {jacobi_sweeps(Lp)}, L^T {jacobi_sweeps(L.T.copy())}, pivoted L^T {jacobi_sweeps(Lp.T.copy())}")
