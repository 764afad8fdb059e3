# scratch: Newton-Schulz polar iteration in the (W, T) form  T_{k+1} = Bm_k T_k  against the current  T_k = W_k K_t,
# emulated split-bf16 products (hi*hi + hi*lo + lo*hi, fp32 accumulate), vs the fp64 SVD.
import math, sys, torch
sys.path.insert(0, "/root/repo/tools/scratch")
from ns_rankdef import COEF, split, q, mm, make, reference, interp_matrix

def run(form, s, t, a, steps=10, rho=None):
    Ns = s.shape[0]
    E = torch.eye(Ns, dtype=torch.float64)
    ref = reference(s, t, a, E)
    s_w, t_w = ref["s_w"], ref["t_w"]
    Kt = q(t_w @ t_w.T).double()
    SW = q(s_w).double()
    W = q(s_w.T).double()
    T = q(mm(W, Kt)).double()
    A = mm(T, W.T).double()
    r = 1.0 / A.diagonal().sum()
    for k in range(steps):
        ca, cb, cc = COEF[k]
        if form == "cur" and k > 0:
            T = q(mm(W, Kt)).double()
        A = mm(T, W.T).double()
        rr = r if k == 0 else 1.0
        As = A * rr
        Bm = q(ca * torch.eye(A.shape[0], dtype=torch.float64) + cb * As + cc * mm(As, As).double()).double()
        sc = math.sqrt(rr)
        Wn = q(sc * mm(Bm, W).double()).double()
        if form == "wt":
            T = q(sc * mm(Bm, T).double()).double()
        W = Wn
    if form == "cur":
        Gsw = mm(Kt, W.T).double()
    else:
        Gsw = T.T            # T = W K_t  =>  T^T = K_t W^T
    nuc = (Gsw * SW).sum()
    Psi = mm(SW, W).double()
    Gt = Psi @ t_w
    e_nuc = abs(nuc - ref["nuc"]) / ref["nuc"]
    e_gs = (Gsw - ref["Gs"]).norm() / ref["Gs"].norm()
    e_gt = (Gt - ref["Gt"]).norm() / ref["Gt"].norm()
    Afin = mm(mm(W, Kt).double(), W.T).double()
    res = (Afin - torch.eye(Afin.shape[0], dtype=torch.float64)).norm()
    return e_nuc.item(), e_gs.item(), e_gt.item(), res.item(), (ref["S"][0] / ref["S"][-1]).item()

for (Ns, Ds, Dt, rho) in [(196, 192, 768, 0.985), (196, 192, 384, 0.985), (196, 192, 768, 0.97), (64, 48, 96, 0.985), (196, 160, 768, 0.96)]:
    s, t, a = make(Ns, Ns, Ds, Dt, r=40, rho=rho)
    for form in ("cur", "wt"):
        e = run(form, s, t, a)
        print(f"N={Ns} Ds={Ds} Dt={Dt} rho={rho} {form:3s}: nuc {e[0]:.2e}  d/ds_w {e[1]:.2e}  d/dt_w {e[2]:.2e}  ||WKW^T-I|| {e[3]:.2e}  cond {e[4]:.1e}")

def run_hybrid(s, t, a, refresh_from, steps=10):
    Ns = s.shape[0]
    E = torch.eye(Ns, dtype=torch.float64)
    ref = reference(s, t, a, E)
    s_w, t_w = ref["s_w"], ref["t_w"]
    Kt = q(t_w @ t_w.T).double(); SW = q(s_w).double(); W = q(s_w.T).double()
    T = q(mm(W, Kt)).double()
    A = mm(T, W.T).double(); r = 1.0 / A.diagonal().sum()
    for k in range(steps):
        ca, cb, cc = COEF[k]
        if k > 0 and k >= refresh_from:
            T = q(mm(W, Kt)).double()
        A = mm(T, W.T).double()
        rr = r if k == 0 else 1.0
        As = A * rr
        Bm = q(ca * torch.eye(A.shape[0], dtype=torch.float64) + cb * As + cc * mm(As, As).double()).double()
        sc = math.sqrt(rr)
        Wn = q(sc * mm(Bm, W).double()).double()
        if k + 1 < refresh_from:
            T = q(sc * mm(Bm, T).double()).double()
        W = Wn
    Gsw = mm(Kt, W.T).double()
    Psi = mm(SW, W).double(); Gt = Psi @ t_w
    return ((Gsw - ref["Gs"]).norm() / ref["Gs"].norm()).item(), ((Gt - ref["Gt"]).norm() / ref["Gt"].norm()).item()

print("hybrid: T updated by Bm until step refresh_from, recomputed from W K_t afterwards")
for (Ns, Ds, Dt, rho) in [(196, 192, 768, 0.985), (196, 192, 384, 0.985), (196, 160, 768, 0.96)]:
    s, t, a = make(Ns, Ns, Ds, Dt, r=40, rho=rho)
    print(Ns, Ds, Dt, rho, [(rf, tuple(f"{x:.1e}" for x in run_hybrid(s, t, a, rf))) for rf in (1, 4, 6, 7, 8, 9, 10)])
