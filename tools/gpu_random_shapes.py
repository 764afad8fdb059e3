"""Random shapes off the BASELINE grid through the drop-in module against the fp32 oracle: MP ranks exact, loss 1e-3, student
gradients 1e-2, temperature gradients 5e-3 of the largest entry (entries that are a cancellation to ~1e-3 of their siblings are
judged against the fp64 oracle by hand, see profiles/r2_random_shapes.txt).  "DEF" marks the feature form with a structurally
rank-deficient teacher token Gram: reported, not counted.  usage: python tools/gpu_random_shapes.py [seed] [count]"""
import sys, os, random, math, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import synth
import test_gpu_parity as T
import vit_bias_aware_structural_distillation_b200 as pkg
pkg.load(); dev = torch.device("cuda:0")
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_bad = 0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    Ds = 8 * rng.randint(2, 30); Dt = 8 * rng.randint(max(Ds // 8, 9), 48)      # (the synthetic single-layer teacher has 64 spikes: D_t >= 72)
    Ns = rng.randint(12, 260); Nt = rng.choice([Ns, Ns, rng.randint(9, 260)])
    has_cls = rng.random() < 0.75
    Lt = rng.randint(1, 5); P = rng.randint(1, 4); H = rng.randint(1, 3) if has_cls else 1
    B = rng.randint(1, 6)
    if B * min(Ns, Nt) < 24: B = 4
    sh = dict(B=B, Ns=Ns, Nt=Nt, Ds=Ds, Dt=Dt, Lt=Lt, H=H, P=P, has_cls=has_cls)
    w = synth.Workload("rand", B, Ns, Nt, Ds, Dt, Lt, H, has_cls, P=P)
    try:
        inp = synth.make_inputs(w, seed=100 + it)
        m = T.build_module(w, dev)
        try:
            ref = T.oracle_case(m, inp, w)
        except Exception as e:                      # torch.linalg raises on the NaN the reference produces at MP rank 0
            print("skip", sh, "reference itself fails:", repr(e)[:80]); continue
        if not math.isfinite(ref["loss"].item()):
            print("skip", sh, "reference itself is not finite (rank 0)"); continue
        deficient = Ds <= min(Ns, Nt) - 1 and (Nt < Ns or Dt < Ns - 1)
        out = T.run_module(m, inp, dev)
        if m.last_polar_residual.item() > m.POLAR_RESIDUAL_OK and not deficient:
            m.polar_steps = 14
            out = T.run_module(m, inp, dev)
        rl = abs(out["loss"].item() - ref["loss"].item()) / abs(ref["loss"].item())
        gt, rt = out["grad_log_temperatures"], ref["grad_log_temperatures"].float()
        tg = ((gt - rt).abs().max() / rt.abs().max().clamp(min=1e-12)).item()
        sg = max(T.rel(out["grad_student"][l], ref["grad_student"][l]) for l in ref["grad_student"])
        ok = out["ranks"] == ref["ranks"] and rl < 1e-3 and sg < 1e-2 and (tg < 5e-3 or Lt == 1)
        if not ok and not deficient: n_bad += 1
        print(("ok  " if ok else ("DEF " if deficient else "BAD ")) + str(sh), f"ranks {'=' if out['ranks'] == ref['ranks'] else str(out['ranks']) + ' vs ' + str(ref['ranks'])} loss {rl:.1e} tgrad {tg:.1e} sgrad {sg:.1e} resid {m.last_polar_residual.item():.1e}", flush=True)
    except Exception as e:
        n_bad += 1
        print("EXC ", sh, repr(e)[:300], flush=True)
print("bad:", n_bad)
