# scratch: feature-form accuracy when the teacher token Gram K_t is rank deficient by construction (N_t < N_s or D_t < N_s - 1)
import sys, os, math, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import synth
import test_gpu_parity as T
import vit_bias_aware_structural_distillation_b200 as pkg
pkg.load(); dev = torch.device("cuda:0")
shapes = [dict(B=4, Ns=251, Nt=122, Ds=104, Dt=376, Lt=1, H=2, P=1, has_cls=True), dict(B=4, Ns=251, Nt=122, Ds=104, Dt=376, Lt=3, H=2, P=2, has_cls=True),
          dict(B=5, Ns=244, Nt=244, Ds=200, Dt=200, Lt=1, H=1, P=1, has_cls=False), dict(B=4, Ns=196, Nt=196, Ds=96, Dt=160, Lt=3, H=2, P=2, has_cls=True),
          dict(B=4, Ns=196, Nt=49, Ds=32, Dt=512, Lt=1, H=1, P=2, has_cls=False), dict(B=4, Ns=196, Nt=100, Ds=64, Dt=256, Lt=3, H=2, P=2, has_cls=True),
          dict(B=4, Ns=256, Nt=196, Ds=192, Dt=768, Lt=3, H=2, P=2, has_cls=True)]
for sh in shapes:
    w = synth.Workload("rand", sh["B"], sh["Ns"], sh["Nt"], sh["Ds"], sh["Dt"], sh["Lt"], sh["H"], sh["has_cls"], P=sh["P"])
    inp = synth.make_inputs(w, seed=7)
    m = T.build_module(w, dev)
    ref = T.oracle_case(m, inp, w, dtype=torch.float64)
    for steps in (8, 9, 10, 12):
        m.polar_steps = steps
        out = T.run_module(m, inp, dev)
        sg = max(T.rel(out["grad_student"][l], ref["grad_student"][l].float()) for l in ref["grad_student"])
        gt, rt = out["grad_log_temperatures"], ref["grad_log_temperatures"].float()
        tg = ((gt - rt).abs().max() / rt.abs().max().clamp(min=1e-12)).item() if sh["Lt"] > 1 else 0.0
        print({k: sh[k] for k in ("Ns", "Nt", "Ds", "Dt", "Lt")}, "steps", steps, f"resid {m.last_polar_residual.item():.2e} loss {abs(out['loss'].item()-ref['loss'].item())/abs(ref['loss'].item()):.1e} tgrad {tg:.1e} sgrad {sg:.1e}", flush=True)
        m._resid_event = None
