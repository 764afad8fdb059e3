"""Round-2 development probe (GPU): shapes beyond one CTA tile / one SM's shared memory against the CPU oracle.
    python tools/gpu_r2_shapes.py [case ...]      cases: eig mp cfg5s cfg4s cfg3s tiny
"""
import ctypes, dataclasses, os, sys, time
import torch, torch.nn as nn
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vit_bias_aware_structural_distillation_b200 as pkg
from oracle import basd_oracle as O, synth

dev = torch.device("cuda:0")
lib = pkg.load()
st = lambda: torch.cuda.current_stream().cuda_stream
rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm().clamp(min=1e-30)).item()

def eig(n):
    torch.manual_seed(n)
    X = torch.randn(4 * n, n) * (0.985 ** torch.arange(n))
    G = X.T @ X
    Gd = G.to(dev)
    ev = torch.zeros(n, device=dev); evec = torch.zeros(n, n, device=dev)
    sw = torch.zeros(4, dtype=torch.int32, device=dev)
    ws = torch.zeros(4 * (3 * n * n + 8 * n) + 16384, dtype=torch.uint8, device=dev)
    t0 = time.time()
    rc = lib.basd_selftest_eig(Gd.data_ptr(), n, ev.data_ptr(), evec.data_ptr(), sw.data_ptr(), ws.data_ptr(), st())
    assert rc == 0, lib.basd_last_error().decode()
    torch.cuda.synchronize()
    dt = time.time() - t0
    ref = torch.linalg.eigvalsh(G.double()).flip(0)
    V = evec.cpu().double()
    print(f"eig n={n}: eval err {((ev.cpu().double() - ref).abs().max() / ref.max()).item():.2e} resid "
          f"{((G.double() @ V.T - V.T * ev.cpu().double()).norm() / G.double().norm()).item():.2e} orth "
          f"{(V @ V.T - torch.eye(n, dtype=torch.float64)).abs().max().item():.2e} sweeps {sw[0].item()} ({dt*1e3:.1f} ms)", flush=True)

def mp(M, D, r):
    g = torch.Generator().manual_seed(M + D)
    f = synth.spiked(1, M, D, r, g)[0]
    t0 = time.time()
    got = pkg.marchenko_pastur_rank(f.to(dev))
    dt = time.time() - t0
    print(f"mp_rank M={M} D={D}: gpu {got} oracle {O.mp_rank(f.float())} ({dt*1e3:.0f} ms)", flush=True)

def run_once(m, inp):
    S = {l: v.to(dev).detach().requires_grad_() for l, v in inp["student"].items()}
    T = {j: v.to(dev) for j, v in inp["teacher"].items()}
    A = {j: v.to(dev) for j, v in inp["attn"].items()}
    logits = inp["logits"].to(dev).requires_grad_()
    m.zero_grad(set_to_none=True)
    loss = m(logits, inp["targets"].to(dev), S, T, A)
    loss.backward()
    torch.cuda.synchronize()
    return loss.detach().clone(), {l: S[l].grad.clone() for l in S}, m.layer_selector.log_temperatures.grad.clone()

def repeat(name, w, seed=1234):
    """two executions of the same step: bitwise equal?"""
    inp = synth.make_inputs(w, seed=seed)
    torch.manual_seed(0)
    m = pkg.BASDLoss(nn.CrossEntropyLoss(label_smoothing=0.001), w.Ds, w.Dt, w.student_depth, w.Ns,
                     config=synth.module_config(w), teacher_has_cls_token=w.has_cls).to(dev)
    a = run_once(m, inp); b = run_once(m, inp)
    same = torch.equal(a[0], b[0]) and torch.equal(a[2], b[2]) and all(torch.equal(a[1][l], b[1][l]) for l in a[1])
    worst = max(rel(a[1][l].float(), b[1][l].float()) for l in a[1])
    print(f"repeat {name}: bitwise {'EQUAL' if same else 'DIFFERENT'} (loss {a[0].item()!r} vs {b[0].item()!r}, worst student-grad rel diff {worst:.2e})", flush=True)

def case(name, w, seed=1234):
    inp = synth.make_inputs(w, seed=seed)
    torch.manual_seed(0)
    m = pkg.BASDLoss(nn.CrossEntropyLoss(label_smoothing=0.001), w.Ds, w.Dt, w.student_depth, w.Ns,
                     config=synth.module_config(w), teacher_has_cls_token=w.has_cls).to(dev)
    S = {l: v.to(dev).detach().requires_grad_() for l, v in inp["student"].items()}
    T = {j: v.to(dev) for j, v in inp["teacher"].items()}
    A = {j: v.to(dev) for j, v in inp["attn"].items()}
    logits = inp["logits"].to(dev).requires_grad_()
    try:
        t0 = time.time()
        loss = m(logits, inp["targets"].to(dev), S, T, A)
        loss.backward()
        torch.cuda.synchronize()
        dt = time.time() - t0
    except Exception as e:
        print(f"{name}: FAILED {type(e).__name__}: {e}", flush=True)
        return
    sel = m.layer_selector
    t0 = time.time()
    ref = O.run_case(inp, sel.proj_s.cpu(), sel.proj_t.cpu(), sel.log_temperatures.detach().cpu(), m.token_layers,
                     has_cls=w.has_cls, n_student_tokens=w.Ns, label_smoothing=0.001)
    dto = time.time() - t0
    gl = {l: rel(S[l].grad.float().cpu(), ref["grad_student"][l]) for l in S}
    gt = sel.log_temperatures.grad.cpu()
    print(f"{name}: loss {loss.item():.6f} ref {ref['loss'].item():.6f} rel {abs(loss.item()-ref['loss'].item())/abs(ref['loss'].item()):.2e} | "
          f"ranks {'OK' if sel.subspace_ranks == ref['ranks'] else str(sel.subspace_ranks) + ' vs ' + str(ref['ranks'])} | "
          f"w err {(sel.last_mixing_weights.cpu() - ref['w'].float()).abs().max().item():.1e} | "
          f"tgrad {gt.tolist()} ref {ref['grad_log_temperatures'].tolist()} | sgrad rel {', '.join(f'{v:.2e}' for v in gl.values())} "
          f"| gpu {dt*1e3:.0f} ms oracle {dto:.1f} s", flush=True)

W = synth.Workload
CASES = {
    "cfg5s": W("cfg5s", 2, 576, 576, 384, 768, 3, 2, True),
    "cfg5m": W("cfg5m", 4, 576, 576, 384, 768, 12, 12, True),
    "cfg4s": W("cfg4s", 4, 196, 196, 384, 1024, 4, 2, True),
    "cfg3s": W("cfg3s", 8, 196, 49, 384, 2048, 1, 1, False),
    "d256": W("d256", 4, 300, 300, 256, 512, 3, 2, True),
    "n320": W("n320", 4, 320, 320, 192, 384, 3, 2, True),
    "cfg1b4": dataclasses.replace(synth.CONFIGS["cfg1"], B=4),
    "cfg2b6": dataclasses.replace(synth.CONFIGS["cfg2"], B=6),
}
IRREG = [dict(B=3, Ns=50, Nt=50, Ds=40, Dt=72, Lt=2, H=3, P=3), dict(B=5, Ns=36, Nt=36, Ds=32, Dt=32, Lt=4, H=1, P=1),
         dict(B=2, Ns=130, Nt=130, Ds=96, Dt=200, Lt=3, H=2, P=2), dict(B=2, Ns=70, Nt=90, Ds=64, Dt=136, Lt=2, H=2, P=4),
         dict(B=9, Ns=210, Nt=210, Ds=200, Dt=256, Lt=2, H=2, P=2), dict(B=3, Ns=160, Nt=160, Ds=144, Dt=192, Lt=2, H=2, P=2),
         dict(B=4, Ns=100, Nt=100, Ds=136, Dt=160, Lt=2, H=2, P=2), dict(B=3, Ns=220, Nt=110, Ds=216, Dt=256, Lt=2, H=2, P=2),
         dict(B=3, Ns=240, Nt=220, Ds=232, Dt=256, Lt=2, H=2, P=2)]
if __name__ == "__main__":
    which = sys.argv[1:] or ["eig", "mp", "n320", "d256", "cfg5s"]
    for c in which:
        if c == "eig":
            for n in (192, 256, 384): eig(n)
        elif c == "mp":
            mp(8000, 384, 30); mp(6000, 768, 40); mp(300, 384, 8)
        elif c == "irreg":
            for k, sh in enumerate(IRREG):
                case(f"irreg{k}", W("irregular", sh["B"], sh["Ns"], sh["Nt"], sh["Ds"], sh["Dt"], sh["Lt"], sh["H"], True, P=sh["P"]), seed=11)
        elif c == "repeat":
            repeat("cfg1b4", CASES["cfg1b4"]); repeat("cfg4s", CASES["cfg4s"]); repeat("cfg2b6", CASES["cfg2b6"]); repeat("cfg5s", CASES["cfg5s"])
        elif c == "tiny":
            from oracle.make_golden import TINY
            for k, w in TINY.items(): case(k, w)
        else:
            case(c, CASES[c])
