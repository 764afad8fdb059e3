"""Generates tests/golden/*.pt by executing the UNMODIFIED reference (/root/reference/src/losses) on the synthetic
inputs of oracle/synth.py.  Run in the build container only (the GPU box has no /root/reference):

    python -m oracle.make_golden [tiny small cfg1 cfg2]

Each fixture stores the workload, seeds and the reference's outputs: loss, MP ranks, log_temperature gradients, and —
because full student gradients of the big configs are too large to commit — their norms, a strided subsample and
inner products with seeded random probes.  The tiny fixtures store full gradients.
"""
from __future__ import annotations

import dataclasses
import os
import sys
import time

import torch
import torch.nn as nn

sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import synth  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

TINY = {
    "tiny_cls": synth.Workload("tiny_cls", 4, 64, 64, 48, 96, 3, 2, True),
    "tiny_interp": synth.Workload("tiny_interp", 4, 64, 36, 48, 96, 3, 2, True),
    "tiny_cnn": synth.Workload("tiny_cnn", 6, 64, 16, 48, 128, 1, 1, False),
    # token-count resampling and CNN (no-CLS) teachers whose resampled token Gram keeps rank >= D_s (the form built so far)
    "tiny_up": synth.Workload("tiny_up", 4, 64, 56, 48, 96, 3, 2, True),            # 56 -> 64 tokens (up-sampling)
    "tiny_down": synth.Workload("tiny_down", 4, 64, 100, 48, 96, 3, 2, True),        # 100 -> 64 tokens (the DINOv2 256 -> 196 case)
    "tiny_cnn_down": synth.Workload("tiny_cnn_down", 6, 64, 81, 48, 128, 1, 1, False),   # 9x9 CNN grid -> 64 student tokens
}


# BASELINE.json configs[2..4] at a reduced batch (the full-size reference step needs 20-100 GB of host RAM and minutes):
# same token counts, widths, layer and head counts - only B differs.
SMALL = {
    "cfg3_b8": dataclasses.replace(synth.CONFIGS["cfg3"], name="cfg3_b8", B=8),      # 7x7 CNN grid -> 196 tokens, D_s = 384 > N_t - 1
    "cfg4_b4": dataclasses.replace(synth.CONFIGS["cfg4"], name="cfg4_b4", B=4),      # D_s = 384 > N - 1 = 195, 24-way mixing
    "cfg5_b2": dataclasses.replace(synth.CONFIGS["cfg5"], name="cfg5_b2", B=2),      # 576 tokens, D_s = 384
}


def probes(shape, n=4, seed=99):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(shape, generator=g) for _ in range(n)]


def run(name: str, w: synth.Workload, label_smoothing=0.001, full=False):
    from src.losses.combined import BASDLoss      # the reference itself
    inp = synth.make_inputs(w)
    torch.manual_seed(0)
    m = BASDLoss(nn.CrossEntropyLoss(label_smoothing=label_smoothing), w.Ds, w.Dt, w.student_depth, w.Ns,
                 config=synth.module_config(w), teacher_has_cls_token=w.has_cls)
    S = {l: v.float().clone().requires_grad_() for l, v in inp["student"].items()}
    T = {j: v.float() for j, v in inp["teacher"].items()}
    logits = inp["logits"].clone().requires_grad_()
    t0 = time.time()
    loss = m(logits, inp["targets"], S, T, inp["attn"])
    loss.backward()
    dt = time.time() - t0
    out = dict(workload=dataclasses.asdict(w), seed=1234, module_seed=0, label_smoothing=label_smoothing,
               loss=loss.detach(), ranks=dict(m.layer_selector.subspace_ranks),
               grad_log_temperatures=m.layer_selector.log_temperatures.grad.clone(), token_layers=list(m.token_layers),
               ref_seconds=dt, torch_version=torch.__version__)
    out["grad_student_norm"] = {l: S[l].grad.norm() for l in S}
    out["grad_student_probe"] = {l: torch.stack([(S[l].grad * p).sum() for p in probes(S[l].shape)]) for l in S}
    out["grad_student_sub"] = {l: S[l].grad.flatten()[::997].clone() for l in S}
    if full:
        out["grad_student"] = {l: S[l].grad.clone() for l in S}
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.save(out, os.path.join(GOLDEN_DIR, f"{name}.pt"))
    print(f"{name}: loss {loss.item():.6f} ranks {out['ranks']} tgrad {out['grad_log_temperatures'].tolist()} ({dt:.1f}s)", flush=True)


def run_standalone():
    """The reference's standalone entry points on their own: geometric_relational_loss (relational.py:5-50), _align_token_count
    (combined.py:9-14) and GrassmannianLayerSelector.forward (layer_selector.py:116-152), values and gradients."""
    from src.losses.combined import _align_token_count
    from src.losses.layer_selector import GrassmannianLayerSelector
    from src.losses.relational import geometric_relational_loss
    w, inp, t_al, attn_same, attn_nocls = synth.standalone_inputs()
    out = dict(torch_version=torch.__version__)
    layer0 = w.token_layers()[0]
    # a11: aligned attention (CLS), attention on another token grid (CLS, 36 -> 48), CNN-style attention without CLS (36 -> 48)
    for key, attn, has_cls in (("pair_cls", attn_same, True), ("pair_cls_resampled", inp["attn"][0].float(), True), ("pair_nocls", attn_nocls, False)):
        s = inp["student"][layer0].float().clone().requires_grad_()
        loss = geometric_relational_loss(s, t_al.float(), attn, has_cls_token=has_cls)
        loss.backward()
        out[key] = dict(loss=loss.detach(), grad_student=s.grad.clone())
    # a8: up- and down-sampling, with the gradient of a seeded linear functional
    for key, n_out in (("align_up", 48), ("align_down", 20), ("align_same", 36)):
        x = inp["teacher"][0].float().clone().requires_grad_()
        y = _align_token_count(x, n_out)
        probe = probes(y.shape, n=1, seed=5)[0]
        (y * probe).sum().backward()
        out[key] = dict(out=y.detach().clone(), grad_in=x.grad.clone(), same_object=y is x)
    # a7: the selector's forward on its own
    torch.manual_seed(0)
    sel = GrassmannianLayerSelector(num_extraction_points=w.P, student_dim=w.Ds, teacher_dim=w.Dt)
    S = {l: v.float().clone().requires_grad_() for l, v in inp["student"].items()}
    T = {j: v.float() for j, v in inp["teacher"].items()}
    A = {j: v.float() for j, v in inp["attn"].items()}
    mixed_t, mixed_a = sel(S, T, A, w.token_layers())
    total = sum((mixed_t[l] * probes(mixed_t[l].shape, n=1, seed=6 + i)[0]).sum() for i, l in enumerate(w.token_layers()))
    total = total + sum((mixed_a[l] * probes(mixed_a[l].shape, n=1, seed=16 + i)[0]).sum() for i, l in enumerate(w.token_layers()))
    total.backward()
    out["selector"] = dict(mixed_tokens={l: mixed_t[l].detach().clone() for l in mixed_t}, mixed_attn_cls_row={l: mixed_a[l][:, :, 0, :].detach().clone() for l in mixed_a},
                           ranks=dict(sel.subspace_ranks), grad_student={l: S[l].grad.clone() for l in S},
                           grad_log_temperatures=sel.log_temperatures.grad.clone())
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.save(out, os.path.join(GOLDEN_DIR, "standalone.pt"))
    print("standalone:", {k: (v["loss"].item() if "loss" in v else "ok") for k, v in out.items() if isinstance(v, dict)}, flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["tiny", "cfg1"]
    for k in which:
        if k == "standalone":
            run_standalone()
        elif k == "tiny":
            for n, w in TINY.items():
                run(n, w, full=True)
        elif k == "small":
            for n, w in SMALL.items():
                run(n, w)
        elif k in SMALL:
            run(k, SMALL[k])
        else:
            run(k, synth.CONFIGS[k])
