# scratch: numerical experiment for rank-deficient Procrustes cores (emulated split-bf16 tensor-core products)
import math, sys, torch
torch.manual_seed(0)
sys.path.insert(0, "/root/repo")
from oracle import basd_oracle as O

COEF = [
    (4.133071044, -11.567471663, 8.093880162),
    (4.132779328, -11.565168373, 8.091968276),
    (4.131589050, -11.555734998, 8.084133962),
    (4.126671962, -11.516778686, 8.051782671),
    (4.106352567, -11.356677293, 7.918925272),
    (4.022478221, -10.711657944, 7.385453943),
    (3.688471560, -8.386451534, 5.490484485),
    (2.745572830, -3.677073572, 1.889197392),
    (1.941173422, -1.383242341, 0.441739912),
    (1.847826056, -1.196240433, 0.348410606),
]

def split(x):
    hi = x.float().bfloat16().float()
    lo = (x.float() - hi).bfloat16().float()
    return hi, lo
def q(x):
    hi, lo = split(x)
    return hi + lo
def mm(A, B, exact=False):
    if exact:
        return (A.double() @ B.double())
    ah, al = split(A); bh, bl = split(B)
    return (ah @ bh + ah @ bl + al @ bh)

def interp_matrix(Ns, Nt):
    E = torch.zeros(Ns, Nt, dtype=torch.float64)
    if Ns == Nt:
        return torch.eye(Ns, dtype=torch.float64)
    for n in range(Ns):
        x = max((n + 0.5) * Nt / Ns - 0.5, 0.0)
        i0 = int(math.floor(x)); i1 = min(i0 + 1, Nt - 1); lam = x - i0
        E[n, i0] += 1 - lam; E[n, i1] += lam
    return E

def make(Ns, Nt, Ds, Dt, r=40, rho=0.985):
    g = torch.Generator().manual_seed(1)
    basis = torch.linalg.qr(torch.randn(Ds, Ds, generator=g))[0]
    s = ((torch.randn(Ns, Ds, generator=g) * 3.0 * rho ** torch.arange(Ds)) @ basis.T).bfloat16().double()
    bt = torch.linalg.qr(torch.randn(Dt, r, generator=g))[0]
    t = ((torch.randn(Nt, r, generator=g) * 4.0 * torch.linspace(1, 0.2, r)) @ bt.T + torch.randn(Nt, Dt, generator=g)).bfloat16().double()
    a = torch.softmax(0.5 * torch.randn(Ns, generator=g), 0).double()
    return s, t, a

def reference(s, t, a, E):
    qv = a.sqrt()
    tal = E @ t
    s_w = qv[:, None] * (s - a @ s)
    t_w = qv[:, None] * (tal - a @ tal)
    C = s_w.T @ t_w
    U, S, Vt = torch.linalg.svd(C, full_matrices=False)
    keep = S > 1e-10 * S[0]
    R = U[:, keep] @ Vt[keep]
    return dict(nuc=S.sum(), Gs=t_w @ R.T, Gt=s_w @ R, s_w=s_w, t_w=t_w, S=S)

def two_sided(s, t, a, E, exact=False, steps=10):
    """V_T: core = teacher token space.  X = L^T Y R, Y (Ns x Nt), right-multiplied."""
    Ns, Nt = E.shape
    qv = a.sqrt()
    one = torch.ones(Nt, dtype=torch.float64) / math.sqrt(Nt)
    Hc = torch.eye(Ns, dtype=torch.float64) - torch.outer(torch.ones(Ns, dtype=torch.float64), a)
    F = qv[:, None] * (Hc @ E)
    s_w = qv[:, None] * (s - a @ s)
    tc = t - t.mean(0)
    KL = (mm(s_w, s_w.T, exact).double() + torch.outer(qv, qv) * 0)   # augmented below
    KR = mm(tc, tc.T, exact).double()
    # augmentation magnitudes: mean eigenvalue scale
    cl = KL.diagonal().sum() / Ns
    cr = KR.diagonal().sum() / Nt
    KL = KL + cl * torch.outer(qv, qv)
    KR = KR + cr * torch.outer(one, one)
    Y0 = F + math.sqrt(1.0) * torch.outer(qv, one)
    # sigma of the decoupled block: sqrt(cl) * 1 * sqrt(cr)
    extra = math.sqrt(cl * cr)
    f = (lambda x: x) if exact else q
    KL, KR, Y = f(KL), f(KR), f(Y0)
    Y0q = Y.clone()
    r = None
    for k in range(steps):
        ca, cb, cc = COEF[k]
        P1 = f(mm(KL, Y, exact))            # Ns x Nt
        P2 = f(mm(KR, Y.T, exact))          # Nt x Ns
        Z = mm(P2, P1, exact)               # Nt x Nt
        if k == 0:
            r = 1.0 / Z.diagonal().sum()
            alpha = math.sqrt(r)
        else:
            r = 1.0
        Z = f(Z * r)
        Bm = f(ca * torch.eye(Nt) + cb * Z + cc * mm(Z, Z, exact))
        Y = f(mm(Y, Bm, exact) * (alpha if k == 0 else 1.0))
    P1 = mm(KL, Y, exact); P2 = mm(KR, Y.T, exact)
    nuc_plus = (P2.T.double() * mm(KL, Y0q, exact).double()).sum()   # tr(K_R Y^T K_L Y_0)
    nuc = nuc_plus - extra
    # G_s = F K_Tc Y^T s_w = F (P2 s_w)   (F 1 = 0 removes the augmentation)
    Gs = F @ mm(P2, s_w, exact).double()
    # G_t (w.r.t. aligned t_w) is not needed; dL/dTbar-side piece: Xi = F^T K_s Y Hp -> compare F^T s_w R == Xi t
    Hp = torch.eye(Nt, dtype=torch.float64) - torch.outer(one, one)
    Xi = F.T @ (P1.double() - cl * torch.outer(qv, qv) @ Y.double()) @ Hp   # F^T q = 0 anyway
    return dict(nuc=nuc, Gs=Gs, Xi=Xi, Y=Y)

def chol_onesided(s, t, a, E, exact=False, steps=10):
    """core = teacher token space through a Cholesky factor: W_0 = G^T F^T (Nt x Ns), K = K_s^+ (Ns x Ns)."""
    Ns, Nt = E.shape
    qv = a.sqrt()
    one = torch.ones(Nt, dtype=torch.float64) / math.sqrt(Nt)
    Hc = torch.eye(Ns, dtype=torch.float64) - torch.outer(torch.ones(Ns, dtype=torch.float64), a)
    F = qv[:, None] * (Hc @ E)
    s_w = qv[:, None] * (s - a @ s)
    tc = t - t.mean(0)
    KL = mm(s_w, s_w.T, exact).double()
    KR = mm(tc, tc.T, exact).double()
    cl = KL.diagonal().sum() / Ns
    cr = KR.diagonal().sum() / Nt
    KL = KL + cl * torch.outer(qv, qv)
    KR = KR + cr * torch.outer(one, one)
    extra = math.sqrt(cl * cr)
    dt = torch.float64 if exact else torch.float32
    G = torch.linalg.cholesky(KR.to(dt)).double()
    Ginv = torch.linalg.inv(G.to(dt)).double() if exact else torch.linalg.solve_triangular(G.float(), torch.eye(Nt), upper=False).double()
    W0 = (G.T @ (F + torch.outer(qv, one)).T)
    if not exact:
        W0 = W0.float().double()
    f = (lambda x: x) if exact else q
    K, W = f(KL), f(W0)
    W0q = W.clone()
    for k in range(steps):
        ca, cb, cc = COEF[k]
        T = f(mm(W, K, exact))
        A = mm(T, W.T, exact)
        if k == 0:
            r = 1.0 / A.diagonal().sum(); alpha = math.sqrt(r)
        else:
            r = 1.0
        A = f(A * r)
        Bm = f(ca * torch.eye(Nt) + cb * A + cc * mm(A, A, exact))
        W = f(mm(Bm, W, exact) * (alpha if k == 0 else 1.0))
    KW = mm(K, W.T, exact).double()           # Ns x Nt
    nuc = (KW * W0q.T.double()).sum() - extra
    X = mm(W, s_w, exact).double()            # Nt x Ds
    Gs = (W0q.T.double() - torch.outer(qv, one) @ G) @ X   # F G X   (W0^T = F G + q 1^T G)
    Gs = (F @ G) @ X
    Xi = F.T @ (KW - cl * torch.outer(qv, qv) @ W.T.double()) @ Ginv @ (torch.eye(Nt, dtype=torch.float64) - torch.outer(one, one))
    return dict(nuc=nuc, Gs=Gs, Xi=Xi)

def rel(a, b):
    return float((a - b).norm() / b.norm())

def run(name, Ns, Nt, Ds, Dt, rho=0.985):
    s, t, a = make(Ns, Nt, Ds, Dt, rho=rho)
    E = interp_matrix(Ns, Nt)
    ref = reference(s, t, a, E)
    qv = a.sqrt()
    Hc = torch.eye(Ns, dtype=torch.float64) - torch.outer(torch.ones(Ns, dtype=torch.float64), a)
    F = qv[:, None] * (Hc @ E)
    Xi_ref_t = F.T @ ref["Gt"]                  # = Xi @ t  (Nt x Dt)
    S = ref["S"]; nz = S[S > 1e-9 * S[0]]
    print(f"== {name}: Ns={Ns} Nt={Nt} Ds={Ds} Dt={Dt} rank={len(nz)} kappa={float(nz[0]/nz[-1]):.3g} nuc={float(ref['nuc']):.6g}")
    for label, fn in (("two-sided", two_sided), ("chol 1-sided", chol_onesided)):
        for exact in (True, False):
            o = fn(s, t, a, E, exact=exact)
            print(f"   {label:13s} {'fp64 ' if exact else 'split'}: nuc rel {abs(float(o['nuc'] - ref['nuc'])) / float(ref['nuc']):.2e}  Gs rel {rel(o['Gs'], ref['Gs']):.2e}"
                  f"  Xi.t rel {rel(o['Xi'] @ t, Xi_ref_t):.2e}")

if __name__ == "__main__" and len(sys.argv) == 1:
    run("cfg4-like", 196, 196, 384, 1024)
    run("cfg3-like", 196, 49, 384, 2048)
    run("tiny-interp-like", 64, 16, 64, 96)
    run("cfg4 steep", 196, 196, 384, 1024, rho=0.97)

def vd_onesided(s, t, a, E, exact=False, steps=10, trace=False):
    """existing V_D: W (Ds x N), K_t (N x N)."""
    Ns, Nt = E.shape
    qv = a.sqrt()
    tal = E @ t
    s_w = qv[:, None] * (s - a @ s)
    t_w = qv[:, None] * (tal - a @ tal)
    f = (lambda x: x) if exact else q
    K = f(mm(t_w, t_w.T, exact).double())
    W = f(s_w.T.clone())
    W0q = W.clone()
    for k in range(steps):
        ca, cb, cc = COEF[k]
        T = f(mm(W, K, exact))
        A = mm(T, W.T, exact)
        if k == 0:
            r = 1.0 / A.diagonal().sum(); alpha = math.sqrt(r)
        else:
            r = 1.0
        A = f(A * r)
        if trace:
            ev = torch.linalg.eigvalsh(((A + A.T) / 2).double())
            print(f"      step {k}: sigma(X) min {float(ev.clamp(min=0).min().sqrt()):.3e} max {float(ev.max().sqrt()):.6f} asym {float((A-A.T).norm()/A.norm()):.2e}")
        Bm = f(ca * torch.eye(A.shape[0]) + cb * A + cc * mm(A, A, exact))
        W = f(mm(Bm, W, exact) * (alpha if k == 0 else 1.0))
    KW = mm(K, W.T, exact).double()
    nuc = (KW * W0q.T.double()).sum()
    return dict(nuc=nuc, Gs=KW)

def run_vd(name, Ns, Nt, Ds, Dt, rho=0.985, trace=False):
    s, t, a = make(Ns, Nt, Ds, Dt, rho=rho)
    E = interp_matrix(Ns, Nt)
    ref = reference(s, t, a, E)
    S = ref["S"]; nz = S[S > 1e-9 * S[0]]
    print(f"== {name}: rank={len(nz)} kappa={float(nz[0]/nz[-1]):.3g} smin/fro={float(nz[-1]/S.norm()):.3g}")
    for exact in (True, False):
        o = vd_onesided(s, t, a, E, exact=exact, trace=trace and not exact)
        print(f"   V_D {'fp64 ' if exact else 'split'}: nuc rel {abs(float(o['nuc'] - ref['nuc'])) / float(ref['nuc']):.2e}  Gs rel {rel(o['Gs'], ref['Gs']):.2e}")

if __name__ == "__main__" and len(sys.argv) == 1:
    run_vd("cfg2-like", 196, 196, 192, 768, trace=True)
    run_vd("cfg2 steep", 196, 196, 192, 768, rho=0.97, trace=True)

def chol_trace(Ns, Nt, Ds, Dt, rho, aug_scale=1.0):
    s, t, a = make(Ns, Nt, Ds, Dt, rho=rho)
    E = interp_matrix(Ns, Nt)
    qv = a.sqrt()
    one = torch.ones(Nt, dtype=torch.float64) / math.sqrt(Nt)
    Hc = torch.eye(Ns, dtype=torch.float64) - torch.outer(torch.ones(Ns, dtype=torch.float64), a)
    F = qv[:, None] * (Hc @ E)
    s_w = qv[:, None] * (s - a @ s)
    tc = t - t.mean(0)
    KL = mm(s_w, s_w.T).double(); KR = mm(tc, tc.T).double()
    print("eig KL", torch.linalg.eigvalsh(KL)[[0,1,2,-1]].tolist())
    print("eig KR", torch.linalg.eigvalsh(KR)[[0,1,2,-1]].tolist())
    cl = KL.diagonal().sum() / Ns * aug_scale; cr = KR.diagonal().sum() / Nt * aug_scale
    KL = KL + cl * torch.outer(qv, qv); KR = KR + cr * torch.outer(one, one)
    G = torch.linalg.cholesky(KR.float()).double()
    W0 = (G.T @ (F + torch.outer(qv, one)).T).float().double()
    K, W = q(KL), q(W0)
    X0 = W.double() @ torch.cat([s_w, qv[:, None] * math.sqrt(cl)], 1)
    sv = torch.linalg.svdvals(X0)
    print("sigma(X0)/fro: max %.4g min %.4g ; aug %.4g" % (float(sv[0] / sv.norm()), float(sv[-1] / sv.norm()), math.sqrt(cl*cr)/float(sv.norm())))
    for k in range(10):
        ca, cb, cc = COEF[k]
        T = q(mm(W, K)); A = mm(T, W.T)
        if k == 0:
            r = 1.0 / A.diagonal().sum(); alpha = math.sqrt(r)
        else:
            r = 1.0
        A = q(A * r)
        ev = torch.linalg.eigvalsh(((A + A.T) / 2).double())
        print(f"      step {k}: sigma(X) min {float(ev.clamp(min=0).min().sqrt()):.3e} max {float(ev.max().sqrt()):.6f} asym {float((A-A.T).norm()/A.norm()):.2e} |W| {float(W.norm()):.3e}")
        Bm = q(ca * torch.eye(A.shape[0]) + cb * A + cc * mm(A, A))
        W = q(mm(Bm, W) * (alpha if k == 0 else 1.0))

if __name__ == "__main__" and len(sys.argv) == 1:
    chol_trace(196, 196, 384, 1024, 0.97)

def steep_study():
    import numpy as np
    for rho in (0.985, 0.96, 0.94, 0.92):
        s, t, a = make(196, 196, 192, 768, rho=rho)
        E = interp_matrix(196, 196)
        ref = reference(s, t, a, E)
        S = ref["S"]
        # fp32 LAPACK reference error (what the reference itself delivers)
        C32 = (ref["s_w"].float().T @ ref["t_w"].float())
        U, Sv, Vt = torch.linalg.svd(C32, full_matrices=False)
        Gs32 = ref["t_w"].float() @ (U @ Vt).T
        print(f"rho={rho}: kappa {float(S[0]/S[-1]):.3g} smin/fro {float(S[-1]/S.norm()):.2e} | fp32 SVD Gs err {rel(Gs32.double(), ref['Gs']):.2e}")
        for extra in (0, 2, 4, 6):
            global COEF
            saved = COEF
            COEF = [saved[0]] * extra + saved
            o = vd_onesided(s, t, a, E, exact=False, steps=10 + extra)
            o64 = vd_onesided(s, t, a, E, exact=True, steps=10 + extra)
            COEF = saved
            print(f"    extra {extra}: split Gs err {rel(o['Gs'], ref['Gs']):.2e} nuc err {abs(float(o['nuc']-ref['nuc']))/float(ref['nuc']):.2e} | fp64-arith Gs err {rel(o64['Gs'], ref['Gs']):.2e}")
if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "steep":
    steep_study()
