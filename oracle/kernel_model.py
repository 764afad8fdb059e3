"""Torch (CPU) model of the ALGORITHM the CUDA kernels implement (TEST INFRASTRUCTURE, not product code).

oracle/basd_oracle.py restates the *reference's* op order (tall SVDs, autograd).  This file restates the
*re-designed* B200 data flow stage by stage — pooled Gram statistics, symmetric eigenproblems instead of
tall SVDs, N-space Procrustes core (Cholesky factor + one-sided Jacobi), closed-form backward — so that
(1) the maths is proven equal to the reference before any kernel is written (tests/test_kernel_model.py
compares it with the oracle/autograd in fp64 and fp32), and (2) every intermediate CUDA buffer has a
plain statement to be compared with in the GPU tests.  Stage names match DESIGN.md / csrc/*.cu.

Reference lines each stage replaces are cited inline (paths under /root/reference/src/losses/).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import basd_oracle as O


def bf16_round(x):
    return x.float().bfloat16().to(x.dtype)


def split_bf16(x):
    hi = bf16_round(x)
    lo = bf16_round(x - hi)
    return hi, lo


# ---------------------------------------------------------------------------------------------
# Stage A: pooled statistics (layer_selector.py:72,13,34-35,86-91 -> Gram + column sums)
# ---------------------------------------------------------------------------------------------
def teacher_stats(Tj, proj_t, dt, emulate_bf16=True):
    """Z = X P_t^T rounded to bf16 (what K2 stores), G = Z^T Z, c = column sums of Z."""
    X = Tj.reshape(-1, Tj.shape[-1]).to(dt)
    P = proj_t.to(dt)
    if emulate_bf16:
        P = bf16_round(P)
    Z = X @ P.T
    if emulate_bf16:
        Z = bf16_round(Z)
    return Z.T @ Z, Z.sum(0), X.shape[0]


def student_stats(Si, dt):
    X = Si.reshape(-1, Si.shape[-1]).to(dt)
    return X.T @ X, X.sum(0), X.shape[0]


# ---------------------------------------------------------------------------------------------
# Stage B: pooled eigenproblems (layer_selector.py:16-19, 36-37, 92)
# ---------------------------------------------------------------------------------------------
def mp_rank_from_gram(G, M, cap):
    """layer_selector.py:8-20 on the uncentred Gram; lower median; strict '>'."""
    D = G.shape[0]
    ev = torch.linalg.eigvalsh(G / M)                  # ascending
    med = ev[(D - 1) // 2]
    lam_plus = med * (1 + math.sqrt(D / M)) ** 2
    return min(int((ev > lam_plus).sum()), cap)


def centred_eig(G, c, M):
    """eigen-decomposition of the centred Gram, DESCENDING order (layer_selector.py:34-37 / :90-92)."""
    Gc = G - torch.outer(c, c) / M
    lam, V = torch.linalg.eigh(Gc)
    return lam.flip(0), V.flip(1)


# ---------------------------------------------------------------------------------------------
# Stage C: principal angles, mixing weights and the pre-computed selector backward
#          (layer_selector.py:95-108)
# ---------------------------------------------------------------------------------------------
def angles_and_gamma(V, lam, Ut_rot, sw, eps32=torch.finfo(torch.float32).eps, want_gamma=True):
    """V [n,n], lam [n] (student, descending, raw basis); Ut_rot = P_s^T U_t [n,k]; sw [k].
    Returns d2 (scalar), cos [k], Gamma_sym [n,n] = d(d2)/dG + transpose (so that d(d2)/dS_c = S_c Gamma_sym)."""
    k = Ut_rot.shape[1]
    n = V.shape[0]
    A = V[:, :k].T @ Ut_rot                                   # k x k
    X, sig, Yt = torch.linalg.svd(A)
    clamp = 1.0 - eps32
    sc = sig.clamp(max=clamp)
    theta = torch.acos(sc)
    d2 = (sw * theta ** 2).sum() / sw.sum()
    if not want_gamma:
        return d2, sig, None
    dsig = sw * 2 * theta * (-1.0 / torch.sqrt(1 - sc ** 2)) / sw.sum()
    dsig = torch.where(sig > clamp, torch.zeros_like(dsig), dsig)
    dA = (X * dsig) @ Yt                                       # X diag(dsig) Y^T
    dVk = Ut_rot @ dA.T                                        # n x k
    Mm = V.T @ dVk                                             # n x k
    Fm = torch.zeros(n, n, dtype=V.dtype)
    denom = lam[:k].unsqueeze(0) - lam[k:].unsqueeze(1)        # [n-k, k]: lam_a - lam_b
    Fm[k:, :k] = Mm[k:, :] / denom
    Gam = V @ Fm @ V.T
    return d2, sig, Gam + Gam.T


def mixing_weights(d2, log_temp):
    tau = F.softplus(log_temp)
    return F.softmax(-d2 / tau, dim=0), tau


def mixing_weights_backward(gw, w, d2, tau, log_temp):
    """returns dL/dd2 [Lt], dL/dlog_temperature."""
    gx = w * (gw - (w * gw).sum())
    gd = -gx / tau
    gtau = (gx * d2).sum() / tau ** 2
    return gd, gtau * torch.sigmoid(log_temp)


# ---------------------------------------------------------------------------------------------
# Stage D: per-sample Procrustes core in token space (relational.py:36-50)
# ---------------------------------------------------------------------------------------------
def procrustes_core(s, tbar, a, factor_side="teacher", Ktt=None, eig_floor=1e-6):
    """One sample.  s [N,Ds], tbar [N,Dt] (mixed + token-aligned teacher), a [N] (sums to 1).

    K_F = L L^T is the (regularised) token Gram of the factor side; the other side o_w enters through
    Y = L^T o_w.  sigma(Y) = sigma(s_w^T t_w).  Returns loss_b, nuc, tr_s, tr_t and the backward pieces
      Gs   [N,Ds] = dL_b/ds          (direct path, unscaled)
      Theta[N,N]  : dL_b/dtbar = 2 (Theta tbar - a mu_t^T)
      ga   [N]    = dL_b/da
    """
    dt = s.dtype
    N = s.shape[0]
    q = a.sqrt()
    mu_s = a @ s
    mu_t = a @ tbar
    s_w = q[:, None] * (s - mu_s)
    if Ktt is None:
        Ktt = tbar @ tbar.T
    m = Ktt @ a
    mm = a @ m
    K_t = q[:, None] * (Ktt - m[:, None] - m[None, :] + mm) * q[None, :]
    K_s_diag = (s_w * s_w).sum(1)
    tr_s = K_s_diag.sum()
    tr_t = K_t.diagonal().sum()
    if factor_side == "teacher":
        KF = K_t
        other = s_w
    else:
        KF = s_w @ s_w.T
        other = q[:, None] * (tbar - mu_t)
    c = KF.diagonal().sum() / N
    L = torch.linalg.cholesky(KF + c * torch.outer(q, q))
    Y = L.T @ other                                            # N x D_other
    if Y.shape[1] <= N:
        # one-sided Jacobi on the columns of Y: Y V = B = U Sigma (no squaring of the condition number)
        U, sig, _ = torch.linalg.svd(Y, full_matrices=False)   # model of the Jacobi result
        keep = sig > eig_floor * sig.max()
    else:
        lamY, U = torch.linalg.eigh(Y @ Y.T)                   # symmetric Jacobi on the N x N Gram
        keep = lamY > eig_floor ** 2 * lamY.max()
        sig = lamY.clamp(min=0).sqrt()
    sig_k = sig[keep]; Uk = U[:, keep]
    nuc = sig_k.sum()
    Omega = (Uk / sig_k) @ Uk.T                                # U Sigma^-1 U^T
    Xi = (Uk * sig_k) @ Uk.T                                   # U Sigma U^T
    Phi = L @ Omega @ L.T                                      # d nuc / d other_w = Phi other_w
    Linv = torch.linalg.inv(L)
    Psi = Linv.T @ Xi @ Linv                                   # d nuc / d factor_w = Psi factor_w
    if factor_side == "teacher":
        Phi_s, Psi_t = Phi, Psi
    else:
        Phi_s, Psi_t = Psi, Phi
    G_sw = Phi_s @ s_w                                         # d nuc / d s_w
    Gs = q[:, None] * (2 * s_w - 2 * G_sw)
    Theta = torch.diag(a) - q[:, None] * Psi_t * q[None, :]
    ga = (K_s_diag + K_t.diagonal() - 2 * (s_w * G_sw).sum(1)) / a
    loss_b = tr_s + tr_t - 2 * nuc
    return dict(loss=loss_b, nuc=nuc, tr_s=tr_s, tr_t=tr_t, Gs=Gs, Theta=Theta, ga=ga, mu_t=mu_t)


# ---------------------------------------------------------------------------------------------
# Whole path: forward + closed-form backward (no autograd anywhere)
# ---------------------------------------------------------------------------------------------
def forward_backward(inputs, proj_s, proj_t, log_temperatures, token_layers, *, has_cls, n_student_tokens,
                     dtype=torch.float64, emulate_bf16=False, ce=None, factor_side=None, allreduce=None, world=1):
    """`allreduce` (in-place sum over ranks) and `world` model the batch-sharded path: every rank
    calls this on its shard; the pooled statistics, d loss / d w and the two UW-SO scalars are summed at exactly the
    points where loss.py issues its collectives.  Returned student gradients follow the CUDA path's convention
    (gradient of the rank-local mean, i.e. world x the global-mean gradient); log_temperature gradients are global."""
    if allreduce is None:
        allreduce = lambda t: t
    dt = dtype
    student = {l: v.to(dt) for l, v in inputs["student"].items()}
    teacher = {j: v.to(dt) for j, v in inputs["teacher"].items()}
    t_idx = sorted(teacher.keys())
    Lt, P = len(t_idx), len(token_layers)
    Ds = proj_s.shape[0]
    Ps = proj_s.to(dt)
    logT = log_temperatures.to(dt)
    B = teacher[t_idx[0]].shape[0]
    Ns = n_student_tokens
    Dt = teacher[t_idx[0]].shape[2]
    if factor_side is None:
        factor_side = "teacher" if Ds <= Ns else "student"

    # Stage A/B teacher
    ranks, Urot, sws = [], [], []
    for j in t_idx:
        G, c, M = teacher_stats(teacher[j], proj_t, dt, emulate_bf16)
        allreduce(G); allreduce(c); M = M * world                     # collective 1 ("stats")
        k = mp_rank_from_gram(G, M, Ds - 1)
        lam, V = centred_eig(G, c, M)
        ranks.append(k)
        Urot.append(Ps.T @ V[:, :k])
        sws.append(lam[:k].clamp(min=0).sqrt())
    rows = torch.stack([O.importance_rows(inputs["attn"][j].to(dt), has_cls) for j in t_idx])   # K1
    rows_i = torch.stack([O.interp_linear_1d(rows[j], Ns) for j in range(Lt)])                # [Lt,B,Ns]
    T_al = torch.stack([O.interp_linear_1d(teacher[j], Ns) for j in t_idx])                    # [Lt,B,Ns,Dt]

    out = dict(ranks=dict(zip(t_idx, ranks)), w=[], d2=[], nuc=[], tr_s=[], tr_t=[], geo_i=[])
    saved = []
    for i, layer in enumerate(token_layers):
        S = student[layer]
        G, c, M = student_stats(S, dt)
        allreduce(G); allreduce(c); M = M * world                     # collective 1 ("stats")
        lam, V = centred_eig(G, c, M)
        d2, gammas = [], []
        for jj in range(Lt):
            d, _, Gam = angles_and_gamma(V, lam, Urot[jj], sws[jj])
            d2.append(d); gammas.append(Gam)
        d2 = torch.stack(d2)
        w, tau = mixing_weights(d2, logT[i])
        # Stage D prep (K6): mix, importance, per sample core
        tbar = (w.view(-1, 1, 1, 1) * T_al).sum(0)
        imp = (w.view(-1, 1, 1) * rows_i).sum(0)
        ssum = imp.sum(-1, keepdim=True)
        a = imp / ssum
        if emulate_bf16:
            th, tl = split_bf16(tbar)
            Ktt_all = th @ th.transpose(1, 2) + th @ tl.transpose(1, 2) + tl @ th.transpose(1, 2)
        cores = []
        for b in range(B):
            Ktt = Ktt_all[b] if emulate_bf16 else None
            cores.append(procrustes_core(S[b], tbar[b], a[b], factor_side, Ktt))
        loss_b = torch.stack([c_["loss"] for c_ in cores])
        out["nuc"].append(torch.stack([c_["nuc"] for c_ in cores]))
        out["tr_s"].append(torch.stack([c_["tr_s"] for c_ in cores]))
        out["tr_t"].append(torch.stack([c_["tr_t"] for c_ in cores]))
        out["geo_i"].append(loss_b.mean()); out["w"].append(w); out["d2"].append(d2)
        saved.append(dict(cores=cores, gammas=gammas, w=w, d2=d2, tau=tau, a=a, ssum=ssum, tbar=tbar, S=S, c=c, M=M))
    geo = torch.stack(out["geo_i"]).mean()
    out["geo"] = geo
    if ce is not None:
        pair = torch.stack([ce.to(dt).detach(), geo.detach()])       # UW-SO weights from the GLOBAL means (loss.py forward)
        allreduce(pair); pair = pair / world
        inv = torch.stack([1.0 / pair[0].clamp(min=torch.finfo(torch.float32).eps),
                           1.0 / pair[1].clamp(min=torch.finfo(torch.float32).eps)])
        omega = inv / inv.sum()
        out["loss"] = omega[0] * ce.to(dt) + omega[1] * geo
        g_geo = omega[1]
    else:
        g_geo = torch.tensor(1.0, dtype=dt)

    # ---- backward (closed form) ----
    grad_student, grad_logT = {}, torch.zeros(P, dtype=dt)
    scale = g_geo / (P * B)
    for i, layer in enumerate(token_layers):
        sv = saved[i]
        Gdir = torch.stack([c_["Gs"] for c_ in sv["cores"]])                     # [B,N,Ds]
        Theta = torch.stack([c_["Theta"] for c_ in sv["cores"]])
        ga = torch.stack([c_["ga"] for c_ in sv["cores"]])
        mu_t = torch.stack([c_["mu_t"] for c_ in sv["cores"]])
        Dt_ = 2 * (Theta @ sv["tbar"] - sv["a"].unsqueeze(-1) * mu_t.unsqueeze(1))   # K10: [B,N,Dt]
        gwt = (ga - (ga * sv["a"]).sum(-1, keepdim=True)) / sv["ssum"]            # d/d w~
        gw = (Dt_.unsqueeze(0) * T_al).sum(dim=(1, 2, 3)) + (gwt.unsqueeze(0) * rows_i).sum(dim=(1, 2))   # K11
        allreduce(gw)                                                  # collective 2 ("gw")
        gw = gw * scale
        gd, glt = mixing_weights_backward(gw, sv["w"], sv["d2"], sv["tau"], logT[i])
        grad_logT[i] = glt / world
        Gam = sum(gd[jj] * sv["gammas"][jj] for jj in range(Lt))                  # K12
        S = sv["S"]
        mu = sv["c"] / sv["M"]
        Sc = S.reshape(-1, Ds) - mu
        grad_student[layer] = scale * Gdir + (Sc @ Gam).reshape(S.shape)           # K13
    out["grad_student"] = grad_student
    out["grad_log_temperatures"] = grad_logT
    out["w"] = torch.stack(out["w"]); out["d2"] = torch.stack(out["d2"])
    out["nuc"] = torch.stack(out["nuc"]); out["tr_s"] = torch.stack(out["tr_s"]); out["tr_t"] = torch.stack(out["tr_t"])
    return out
