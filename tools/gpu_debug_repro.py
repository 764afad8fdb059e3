"""Development aid: run the four phases twice on identical inputs and report where run-to-run differences
(split-K / reduction atomics) enter and how much they are amplified."""
import ctypes, dataclasses, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn as nn
import __graft_entry__ as g
g.build()
import vit_bias_aware_structural_distillation_b200 as pkg
from vit_bias_aware_structural_distillation_b200 import _lib, loss as L
from oracle import synth
lib = pkg.load(); dev = torch.device("cuda:0")
def st(): return torch.cuda.current_stream().cuda_stream
w = dataclasses.replace(synth.CONFIGS["cfg2"], B=6)
inp = synth.make_inputs(w)
torch.manual_seed(0)
m = pkg.BASDLoss(nn.CrossEntropyLoss(label_smoothing=0.001), w.Ds, w.Dt, w.student_depth, w.Ns, config=synth.module_config(w), teacher_has_cls_token=w.has_cls).to(dev)
sel = m.layer_selector
students = [inp["student"][l].to(dev) for l in m.token_layers]; teachers = [inp["teacher"][j].to(dev) for j in sorted(inp["teacher"])]
attns = [inp["attn"][j].to(dev) for j in sorted(inp["attn"])]
shape, cin, keep = L._prepare(students, teachers, attns, sel.proj_s, sel.proj_t, sel.log_temperatures, w.has_cls, 1)
nb = ctypes.c_size_t(); _lib.check(lib.basd_workspace_bytes(ctypes.byref(shape), ctypes.byref(nb)), "ws")
names = ["stats", "evals", "evecs", "d2", "w", "gamma", "a", "ktt", "polar_gsw", "gdir", "gwt", "loss_b", "gw", "corr"]
def once():
    ws = torch.zeros(nb.value, dtype=torch.uint8, device=dev); geo = torch.zeros((), device=dev)
    one = torch.ones((), device=dev)
    _lib.check(lib.basd_forward_stats(ctypes.byref(shape), ctypes.byref(cin), ws.data_ptr(), st()), "stats")
    _lib.check(lib.basd_forward_solve(ctypes.byref(shape), ctypes.byref(cin), ws.data_ptr(), geo.data_ptr(), st()), "solve")
    _lib.check(lib.basd_backward_dots(ctypes.byref(shape), ctypes.byref(cin), ws.data_ptr(), st()), "dots")
    grads = [torch.empty(s_.shape, dtype=torch.float32, device=dev) for s_ in students]
    glt = torch.empty(w.P, device=dev)
    ptrs = (ctypes.c_void_p * len(grads))(*[t.data_ptr() for t in grads])
    _lib.check(lib.basd_backward_finish(ctypes.byref(shape), ctypes.byref(cin), ws.data_ptr(), one.data_ptr(), ptrs, 0, glt.data_ptr(), st()), "finish")
    torch.cuda.synchronize()
    out = {n: L.workspace_view(shape, ws, n).clone() for n in names}
    out["grad0"] = grads[0]; out["grad3"] = grads[3]; out["glt"] = glt; out["geo"] = geo.view(1)
    return out
a, b = once(), once()
for n in a:
    x, y = a[n].double(), b[n].double()
    d = (x - y).norm() / y.norm().clamp(min=1e-300)
    print(f"{n:10s} rel diff between two runs {d.item():.3e}   (norm {y.norm().item():.3e})")
# selector share of the student gradient
print("DONE")
