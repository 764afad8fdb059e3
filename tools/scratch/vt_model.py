# fp64 model of the V_T pipeline exactly as the kernels will run it; checked against autograd of the reference formula
import math, sys, torch
sys.path.insert(0, "/root/repo")
from tools.scratch.ns_rankdef import interp_matrix, make, COEF
torch.set_default_dtype(torch.float64)

def ref_loss(s, t, a, E):
    qv = a.sqrt(); tal = E @ t
    s_w = qv[:, None] * (s - a @ s); t_w = qv[:, None] * (tal - a @ tal)
    return (s_w ** 2).sum() + (t_w ** 2).sum() - 2 * torch.linalg.matrix_norm(s_w.T @ t_w, ord="nuc")

def vt(s, t, a, E, steps=10):
    Ns, Nt = E.shape; Ds = s.shape[1]
    qv = a.sqrt()
    s_w = qv[:, None] * (s - a @ s)
    Ktt = t @ t.T
    rm = Ktt.mean(1); tot = rm.mean()
    K = Ktt - rm[:, None] - rm[None, :] + tot
    cr = K.diagonal().sum() / Nt
    K = K + cr / Nt
    G = torch.linalg.cholesky(K)
    Ginv = torch.linalg.inv(G)
    ebar = E.T @ a
    eg = ebar @ G
    FG = qv[:, None] * (E @ G - eg[None, :])
    ktd = (FG ** 2).sum(1); tr_t = ktd.sum()
    X0 = FG.T @ s_w
    fro2 = (X0 ** 2).sum()
    beta = math.sqrt(fro2 / (Nt - 1))
    zhat = G.sum(0) / math.sqrt(Nt) / math.sqrt(cr)
    Ds16 = (Ds + 15) // 16 * 16
    Xp = torch.zeros(Nt, Ds16 + 8); Xp[:, :Ds] = X0; Xp[:, Ds16] = beta * zhat
    X0p = Xp.clone()
    X = Xp
    for k in range(steps):
        ca, cb, cc = COEF[k]
        A = X @ X.T
        r = 1.0 / A.diagonal().sum() if k == 0 else 1.0
        A = A * r
        Bm = ca * torch.eye(Nt) + cb * A + cc * A @ A
        X = (math.sqrt(r) if k == 0 else 1.0) * Bm @ X
    Gsw = FG @ X[:, :Ds]
    dots = (s_w * Gsw).sum(1)
    nuc = dots.sum()
    loss = (s_w ** 2).sum() + tr_t - 2 * nuc
    H = X0p[:, :Ds16] @ X[:, :Ds16].T
    GinvC = Ginv - Ginv.mean(1, keepdim=True)
    Thraw = Ginv.T @ (H @ GinvC)
    FtF = (E.T * a) @ E - torch.outer(ebar, ebar)
    Theta = 2 * FtF - 2 * Thraw
    gdir = qv[:, None] * (2 * s_w - 2 * Gsw)
    ksd = (s_w ** 2).sum(1)
    ga = (ksd + ktd - 2 * dots) / a
    return dict(loss=loss, dT=Theta @ t, gdir=gdir, ga=ga, aug_sigma=(X[:, Ds16] ** 2).sum().sqrt())

for (Ns, Nt, Ds, Dt) in [(64, 36, 48, 96), (64, 16, 48, 128), (196, 49, 384, 512), (96, 96, 128, 160)]:
    s, t, a = make(Ns, Nt, Ds, Dt)
    E = interp_matrix(Ns, Nt)
    s = s.clone().requires_grad_(); t = t.clone().requires_grad_(); a_ = a.clone().requires_grad_()
    L = ref_loss(s, t, a_, E); L.backward()
    o = vt(s.detach(), t.detach(), a.detach(), E)
    rel = lambda x, y: float((x - y).norm() / y.norm())
    print(f"Ns={Ns} Nt={Nt} Ds={Ds}: loss rel {abs(float(o['loss'] - L)) / abs(float(L)):.2e} dT rel {rel(o['dT'], t.grad):.2e} "
          f"gdir rel {rel(o['gdir'], s.grad):.2e} ga rel {rel(o['ga'], a_.grad):.2e} aug sigma {float(o['aug_sigma']):.6f}")
