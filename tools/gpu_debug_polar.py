"""Development aid: checks the Newton-Schulz polar path stage by stage against torch (fp64) on the GPU box."""
import ctypes, dataclasses, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn as nn
import __graft_entry__ as g
g.build()
import vit_bias_aware_structural_distillation_b200 as pkg
from vit_bias_aware_structural_distillation_b200 import _lib, loss as L
from oracle import synth, basd_oracle as O
lib = pkg.load(); dev = torch.device("cuda:0")
def st(): return torch.cuda.current_stream().cuda_stream

def run(name, B):
    w = dataclasses.replace(synth.CONFIGS[name], B=B)
    inp = synth.make_inputs(w)
    torch.manual_seed(0)
    m = pkg.BASDLoss(nn.CrossEntropyLoss(label_smoothing=0.001), w.Ds, w.Dt, w.student_depth, w.Ns, config=synth.module_config(w), teacher_has_cls_token=w.has_cls).to(dev)
    sel = m.layer_selector
    students = [inp["student"][l].to(dev) for l in m.token_layers]; teachers = [inp["teacher"][j].to(dev) for j in sorted(inp["teacher"])]
    attns = [inp["attn"][j].to(dev) for j in sorted(inp["attn"])]
    shape, cin, keep = L._prepare(students, teachers, attns, sel.proj_s, sel.proj_t, sel.log_temperatures, w.has_cls, 1)
    nb = ctypes.c_size_t(); _lib.check(lib.basd_workspace_bytes(ctypes.byref(shape), ctypes.byref(nb)), "ws")
    ws = torch.zeros(nb.value, dtype=torch.uint8, device=dev); geo = torch.zeros((), device=dev)
    _lib.check(lib.basd_forward_stats(ctypes.byref(shape), ctypes.byref(cin), ws.data_ptr(), st()), "stats")
    _lib.check(lib.basd_forward_solve(ctypes.byref(shape), ctypes.byref(cin), ws.data_ptr(), geo.data_ptr(), st()), "solve")
    torch.cuda.synchronize()
    V = lambda n, dt=torch.float32: L.workspace_view(shape, ws, n, dt)
    P, N, Ds, Np = w.P, w.Ns, w.Ds, (w.Ns + 63) // 64 * 64
    nprob = P * B
    def split(name, rows, pitch, inner):        # column-block tiled storage [2][prob][col block][row][64]
        cb = (inner + 63) // 64
        x = V(name, torch.bfloat16).view(2, nprob, cb, rows, 64).double()
        x = (x[0] + x[1]).permute(0, 2, 1, 3).reshape(nprob, rows, cb * 64)
        return x[..., :inner]
    a = V("a").view(nprob, N).double(); q = a.sqrt()
    ktt = V("ktt").view(nprob, N, N).double()
    mvec = (ktt @ a.unsqueeze(-1)).squeeze(-1); mm = (a * mvec).sum(-1)
    Kt_ref = q.unsqueeze(-1) * (ktt - mvec.unsqueeze(-1) - mvec.unsqueeze(-2) + mm.view(-1, 1, 1)) * q.unsqueeze(-2)
    Kt = split("polar_kt", N, Np, N)
    print(f"  K_t rel err {((Kt - Kt_ref).norm() / Kt_ref.norm()).item():.2e}")
    S = torch.stack([s_.double() for s_ in students]).view(nprob, N, Ds)
    mu = (a.unsqueeze(-1) * S).sum(1, keepdim=True)
    sw_ref = q.unsqueeze(-1) * (S - mu)
    SW = split("polar_sw", N, Ds, Ds)
    print(f"  s_w rel err {((SW - sw_ref).norm() / sw_ref.norm()).item():.2e}")
    C_fro2 = torch.einsum("pnd,pnm,pme->pde", sw_ref, Kt_ref, sw_ref).diagonal(dim1=1, dim2=2).sum(-1)
    fro2 = V("polar_fro2").double()
    print(f"  ||C||_F^2 rel err {((fro2 - C_fro2).abs() / C_fro2).max().item():.2e}")
    W = split("polar_w", Ds, Np, N)                       # [p][Ds][N]
    # X = W t_w ; X X^T = W K_t W^T should be the identity
    A = W @ Kt_ref @ W.transpose(1, 2)
    eye = torch.eye(Ds, dtype=torch.float64, device=dev)
    print(f"  ||W K_t W^T - I||_max {(A - eye).abs().max().item():.2e}")
    # truth from the SVD of the weighted cross-covariance through the Cholesky-free route: nuc = sum sqrt(eig(sw^T Kt sw))
    M = sw_ref.transpose(1, 2) @ Kt_ref @ sw_ref
    nuc_ref = torch.linalg.eigvalsh(M).clamp(min=0).sqrt().sum(-1)
    dbg = V("dbg").view(nprob, 5).double()
    print(f"  nuc rel err {((dbg[:, 0] - nuc_ref).abs() / nuc_ref).max().item():.2e}")
    lam, U = torch.linalg.eigh(M)
    Minv_half = (U / lam.clamp(min=1e-300).sqrt().unsqueeze(-2)) @ U.transpose(1, 2)
    G_ref = Kt_ref @ sw_ref @ Minv_half                  # K_t s_w (C C^T)^-1/2 = t_w R^T
    G = V("polar_gsw").view(nprob, N, Ds).double()
    print(f"  Gsw rel err {((G - G_ref).norm() / G_ref.norm()).item():.2e}")
    print(f"  geo {geo.item():.6f}")
    m2 = pkg.BASDLoss(nn.CrossEntropyLoss(label_smoothing=0.001), w.Ds, w.Dt, w.student_depth, w.Ns, config=synth.module_config(w), teacher_has_cls_token=w.has_cls).to(dev)
    m2.load_state_dict(m.state_dict())
    ref = O.run_case(inp, sel.proj_s.cpu(), sel.proj_t.cpu(), sel.log_temperatures.detach().cpu(), m.token_layers, has_cls=w.has_cls, n_student_tokens=w.Ns, label_smoothing=0.001)
    print(f"  geo ref {ref['geo'].item():.6f} rel {abs(geo.item() - ref['geo'].item()) / ref['geo'].item():.2e}")
    Sg = {l: v.to(dev).requires_grad_() for l, v in inp["student"].items()}
    T = {j: v.to(dev) for j, v in inp["teacher"].items()}; A_ = {j: v.to(dev) for j, v in inp["attn"].items()}
    logits = inp["logits"].to(dev).requires_grad_()
    for rep in range(3):
        for l in Sg: Sg[l].grad = None
        m2.zero_grad(set_to_none=True)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
        e0.record(); loss = m2(logits, inp["targets"].to(dev), Sg, T, A_); e1.record(); loss.backward(); e2.record(); torch.cuda.synchronize()
    print(f"  timed fwd {e0.elapsed_time(e1):.2f} ms bwd {e1.elapsed_time(e2):.2f} ms; loss {loss.item():.6f} ref {ref['loss'].item():.6f}")
    gt = m2.layer_selector.log_temperatures.grad.cpu()
    print(f"  tgrad rel {((gt - ref['grad_log_temperatures']).abs() / ref['grad_log_temperatures'].abs()).max().item():.2e}")
    for l in m.token_layers:
        gg = Sg[l].grad.float().cpu(); rg = ref["grad_student"][l]
        print(f"  layer {l}: student grad rel {((gg - rg).norm() / rg.norm()).item():.3e}")

for name, B in [("cfg1", 2), ("cfg2", 8)]:
    print(f"===== {name} B={B}", flush=True)
    try:
        run(name, B)
    except Exception:
        import traceback; traceback.print_exc()
print("DONE")
