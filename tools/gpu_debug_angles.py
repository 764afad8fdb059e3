"""Development aid: phase clocks of one angles CTA (student point 0, teacher layer BASD_SPECTRAL_DBG - 1) inside a cfg2 step."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("BASD_SPECTRAL_DBG", "12")
import torch, torch.nn as nn
import bench
import vit_bias_aware_structural_distillation_b200 as pkg
from oracle import synth
lib = pkg.load(); dev = torch.device("cuda:0")
w = bench.workload(256, sys.argv[1] if len(sys.argv) > 1 else "cfg2")
logits, targets, student, teacher, attn = bench.device_inputs(w, dev, 1)
torch.manual_seed(0)
m = pkg.BASDLoss(nn.CrossEntropyLoss(), w.Ds, w.Dt, w.student_depth, w.Ns, config=synth.module_config(w), teacher_has_cls_token=w.has_cls).to(dev)
for _ in range(3): m.geo_loss(student, teacher, attn)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 32)(); lib.basd_debug_spectral_clocks(buf); c = list(buf)
names = ["Ur", "W", "-", "jacobi", "distances", "Q", "WQ", "F", "H", "Gamma_sym"]
print("ranks", dict(m.layer_selector.subspace_ranks))
print("pooled_eig: pre", c[0], "jacobi", c[1], "post", c[2], "sweeps", c[3])
print(f"angles CTA (layer {int(os.environ['BASD_SPECTRAL_DBG']) - 1}, point 0): k =", c[31], " ".join(f"{n}={c[8+i+1]-c[8+i]}" for i, n in enumerate(names)),
      "total", c[8 + 10] - c[8], "jacobi sweeps", c[30])
