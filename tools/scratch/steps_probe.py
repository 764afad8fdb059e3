# scratch: accuracy and residual of the polar iteration against the number of Newton-Schulz steps (cfg1 shapes, B = 32; cfg2 shapes, B = 16)
import sys, os, dataclasses, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import synth
import test_gpu_parity as T
import vit_bias_aware_structural_distillation_b200 as pkg
pkg.load(); dev = torch.device("cuda:0")
for name, B in (("cfg1", 32), ("cfg2", 16)):
    w = dataclasses.replace(synth.CONFIGS[name], B=B)
    inp = synth.make_inputs(w)
    m = T.build_module(w, dev)
    ref = T.oracle_case(m, inp, w)
    for steps in (10, 9, 8, 7):
        m.polar_steps = steps
        out = T.run_module(m, inp, dev)
        gt, rt = out["grad_log_temperatures"], ref["grad_log_temperatures"].float()
        sg = max(T.rel(out["grad_student"][l], ref["grad_student"][l]) for l in ref["grad_student"])
        print(f"{name} B={B} steps={steps}: residual {m.last_polar_residual.item():.3e} loss rel {abs(out['loss'].item()-ref['loss'].item())/abs(ref['loss'].item()):.2e} "
              f"tgrad rel {((gt-rt).abs()/rt.abs()).max().item():.2e} sgrad rel {sg:.2e}", flush=True)
