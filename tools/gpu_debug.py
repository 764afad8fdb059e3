"""Stage-by-stage GPU diagnostics (development aid; prints instead of asserting)."""
import ctypes, dataclasses, math, os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn as nn
import __graft_entry__ as g
g.build()
import vit_bias_aware_structural_distillation_b200 as pkg
from vit_bias_aware_structural_distillation_b200 import _lib, loss as L
from oracle import synth, basd_oracle as O, kernel_model as K
lib = pkg.load(); dev = torch.device("cuda:0")
torch.manual_seed(1)
def st(): return torch.cuda.current_stream().cuda_stream
def section(name):
    print(f"\n===== {name}", flush=True)

def gemm_case(variant, M, N, K_):
    gen = torch.Generator(device="cpu").manual_seed(variant * 1000 + M + N + K_)
    if variant == 0:
        A = torch.randn(M, K_, generator=gen).bfloat16(); B = torch.randn(N, K_, generator=gen).bfloat16(); ref = A.float() @ B.float().T
    elif variant == 1:
        A = torch.randn(K_, M, generator=gen).bfloat16(); B = torch.randn(K_, N, generator=gen).bfloat16(); ref = A.float().T @ B.float()
    elif variant == 2:
        A = torch.randn(M, K_, generator=gen).bfloat16(); B = torch.randn(K_, N, generator=gen).bfloat16(); ref = A.float() @ B.float()
    else:
        X = torch.randn(M, K_, generator=gen); A = X.bfloat16(); B = (X - A.float()).bfloat16(); N = M
        ref = A.float() @ A.float().T + A.float() @ B.float().T + B.float() @ A.float().T
    Ad, Bd = A.to(dev), B.to(dev); C = torch.zeros(M, N, device=dev)
    rc = lib.basd_selftest_gemm(variant, Ad.data_ptr(), Bd.data_ptr(), C.data_ptr(), M, N, K_, st())
    if rc: print(f"  gemm v{variant} {M}x{N}x{K_}: ERROR {lib.basd_last_error().decode()}"); return
    torch.cuda.synchronize()
    err = (C.cpu() - ref).abs().max().item() / ref.abs().max().item()
    print(f"  gemm v{variant} {M}x{N}x{K_}: max rel err {err:.3e} {'OK' if err < 1e-4 else 'FAIL'}", flush=True)

section("tcgen05 GEMM self tests")
for (v, M, N, K_) in [(0, 128, 192, 64), (0, 256, 192, 128), (0, 300, 200, 384), (0, 1000, 384, 768),
                      (1, 192, 192, 256), (1, 192, 192, 1000), (1, 384, 384, 640), (1, 48, 48, 256),
                      (2, 196, 128, 200), (2, 196, 768, 200), (2, 64, 96, 64),
                      (3, 196, 196, 768), (3, 64, 64, 96), (3, 200, 200, 128)]:
    try: gemm_case(v, M, N, K_)
    except Exception: traceback.print_exc()

section("Jacobi eigen self test")
for n in (48, 192):
    try:
        X = torch.randn(4 * n, n) * (0.97 ** torch.arange(n)); G = (X.T @ X)
        Gd = G.to(dev); ev = torch.zeros(n, device=dev); evec = torch.zeros(n, n, device=dev); sw = torch.zeros(4, dtype=torch.int32, device=dev)
        ws = torch.zeros(4 * (2 * n * n + n) + 8192, dtype=torch.uint8, device=dev)
        t0 = time.time()
        rc = lib.basd_selftest_eig(Gd.data_ptr(), n, ev.data_ptr(), evec.data_ptr(), sw.data_ptr(), ws.data_ptr(), st())
        torch.cuda.synchronize(); dt = time.time() - t0
        if rc: print("  eig ERROR", lib.basd_last_error().decode()); continue
        ref = torch.linalg.eigvalsh(G.double()).flip(0)
        V = evec.cpu().double()
        res = (G.double() @ V.T - V.T * ev.cpu().double()).norm() / G.double().norm()
        orth = (V @ V.T - torch.eye(n, dtype=torch.float64)).abs().max()
        print(f"  eig n={n}: sweeps {sw[0].item()} time {dt*1e3:.2f} ms  eval rel err {((ev.cpu().double()-ref).abs().max()/ref.max()).item():.2e} residual {res.item():.2e} orth {orth.item():.2e}", flush=True)
    except Exception: traceback.print_exc()

def run_cfg(name, B, check_stages=True):
    section(f"full path {name} B={B}")
    w = dataclasses.replace(synth.CONFIGS[name], B=B) if name in synth.CONFIGS else name
    inp = synth.make_inputs(w)
    torch.manual_seed(0)
    m = pkg.BASDLoss(nn.CrossEntropyLoss(label_smoothing=0.001), w.Ds, w.Dt, w.student_depth, w.Ns, config=synth.module_config(w), teacher_has_cls_token=w.has_cls).to(dev)
    sel = m.layer_selector
    S = {l: v.to(dev).requires_grad_() for l, v in inp["student"].items()}
    T = {j: v.to(dev) for j, v in inp["teacher"].items()}; A = {j: v.to(dev) for j, v in inp["attn"].items()}
    logits = inp["logits"].to(dev).requires_grad_()
    ref = O.run_case(inp, sel.proj_s.cpu(), sel.proj_t.cpu(), sel.log_temperatures.detach().cpu(), m.token_layers, has_cls=w.has_cls, n_student_tokens=w.Ns, label_smoothing=0.001)
    # raw phases for stage checks
    students = [S[l].detach() for l in m.token_layers]; teachers = [T[j] for j in sorted(T)]; attns = [A[j] for j in sorted(A)]
    shape, cin, keep = L._prepare(students, teachers, attns, sel.proj_s, sel.proj_t, sel.log_temperatures, w.has_cls, 1)
    nb = ctypes.c_size_t(); _lib.check(lib.basd_workspace_bytes(ctypes.byref(shape), ctypes.byref(nb)), "ws")
    print(f"  workspace {nb.value/1e6:.1f} MB")
    ws = torch.zeros(nb.value, dtype=torch.uint8, device=dev); geo = torch.zeros((), device=dev)
    _lib.check(lib.basd_forward_stats(ctypes.byref(shape), ctypes.byref(cin), ws.data_ptr(), st()), "stats"); torch.cuda.synchronize()
    V = lambda n, dt=torch.float32: L.workspace_view(shape, ws, n, dt)
    if check_stages:
        rows_ref = torch.stack([O.importance_rows(inp["attn"][j].float(), w.has_cls) for j in sorted(inp["attn"])])
        print(f"  rows max abs err {(V('rows').cpu().view_as(rows_ref)-rows_ref).abs().max().item():.2e}")
        n = w.Ds; stats = V("stats").cpu().view(w.Lt + w.P, n * n + n)
        for idx in (0, w.Lt - 1):
            Gm, c, M = K.teacher_stats(inp["teacher"][idx].float(), sel.proj_t.cpu(), torch.float32, True)
            # model uses single-bf16 P; kernel uses split P: compare against split version
            X = inp["teacher"][idx].float().reshape(-1, w.Dt); hi, lo = K.split_bf16(sel.proj_t.cpu()); Z = K.bf16_round(X @ hi.T + X @ lo.T)
            Gs = Z.T @ Z; cs = Z.sum(0)
            print(f"  teacher {idx}: gram rel err {((stats[idx,:n*n].view(n,n)-Gs).abs().max()/Gs.abs().max()).item():.2e} colsum rel err {((stats[idx,n*n:]-cs).abs().max()/cs.abs().max()).item():.2e}")
        for i, l in enumerate(m.token_layers[:2]):
            Gs, cs, M = K.student_stats(inp["student"][l].float(), torch.float32)
            print(f"  student {i}: gram rel err {((stats[w.Lt+i,:n*n].view(n,n)-Gs).abs().max()/Gs.abs().max()).item():.2e} colsum rel err {((stats[w.Lt+i,n*n:]-cs).abs().max()/cs.abs().max()).item():.2e}")
    t0 = time.time()
    _lib.check(lib.basd_forward_solve(ctypes.byref(shape), ctypes.byref(cin), ws.data_ptr(), geo.data_ptr(), st()), "solve"); torch.cuda.synchronize()
    print(f"  forward_solve {1e3*(time.time()-t0):.1f} ms; sweeps {V('sweeps', torch.int32).tolist()}")
    ranks = V("ranks", torch.int32).tolist()
    print(f"  ranks {ranks} ref {list(ref['ranks'].values())} {'OK' if ranks == list(ref['ranks'].values()) else 'MISMATCH'}")
    d2 = V("d2").cpu().view(w.P, w.Lt); wv = V("w").cpu().view(w.P, w.Lt)
    print(f"  d2 rel err {((d2-ref['d2']).abs().max()/ref['d2'].abs().max()).item():.2e}  w max abs err {(wv-ref['w']).abs().max().item():.2e}")
    dbg = V("dbg").cpu().view(w.P, B, 5)
    print(f"  nuc rel err {((dbg[...,0]-ref['nuc']).abs()/ref['nuc']).max().item():.2e} tr_s rel {((dbg[...,1]-ref['tr_s']).abs()/ref['tr_s']).max().item():.2e} tr_t rel {((dbg[...,2]-ref['tr_t']).abs()/ref['tr_t']).max().item():.2e} jacobi sweeps {dbg[...,3].min().item():.0f}-{dbg[...,3].max().item():.0f} chol_bad {dbg[...,4].max().item():.0f}")
    print(f"  geo {geo.item():.6f} ref {ref['geo'].item():.6f} rel {abs(geo.item()-ref['geo'].item())/ref['geo'].item():.2e}")
    del ws
    # module path with autograd
    t0 = time.time(); loss = m(logits, inp["targets"].to(dev), S, T, A); loss.backward(); torch.cuda.synchronize()
    print(f"  module fwd+bwd wall {1e3*(time.time()-t0):.1f} ms (first call)")
    for rep in range(2):
        for l in S: S[l].grad = None
        sel.log_temperatures.grad = None
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
        e0.record(); loss = m(logits, inp["targets"].to(dev), S, T, A); e1.record(); loss.backward(); e2.record(); torch.cuda.synchronize()
        print(f"  timed: fwd {e0.elapsed_time(e1):.2f} ms bwd {e1.elapsed_time(e2):.2f} ms")
    print(f"  loss {loss.item():.6f} ref {ref['loss'].item():.6f} rel {abs(loss.item()-ref['loss'].item())/ref['loss'].item():.2e}")
    gt = sel.log_temperatures.grad.cpu()
    print(f"  tgrad {gt.tolist()} ref {ref['grad_log_temperatures'].tolist()} rel {((gt-ref['grad_log_temperatures']).abs()/ref['grad_log_temperatures'].abs()).max().item():.2e}")
    for l in m.token_layers:
        gg = S[l].grad.float().cpu(); rg = ref["grad_student"][l]
        print(f"  layer {l}: student grad rel {((gg-rg).norm()/rg.norm()).item():.3e}")
    # direct-only comparison
    refd = O.run_case(inp, sel.proj_s.cpu(), sel.proj_t.cpu(), sel.log_temperatures.detach().cpu(), m.token_layers, has_cls=w.has_cls, n_student_tokens=w.Ns, label_smoothing=0.001, detach_weights=True)
    return m, S, T, A, logits, inp

for name, B in [("cfg1", 4), ("cfg1", 32), ("cfg2", 16)]:
    try: run_cfg(name, B)
    except Exception: traceback.print_exc()
print("DONE", flush=True)
