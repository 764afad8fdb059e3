"""Synthetic BASD activations (TEST INFRASTRUCTURE — not product code).

Implements the generator of SURVEY.md Appendix D: "spiked" teacher tokens, geometric-spectrum
student tokens, softmaxed random attention.  i.i.d. Gaussian teacher tokens make the reference
return NaN (Marchenko-Pastur rank 0 -> 0/0 at /root/reference/src/losses/layer_selector.py:105),
so every parity input is spiked.  All values are rounded to bf16 once and handed to the fp32
oracle as `.float()` of the same numbers, so the CUDA path and the oracle see identical inputs.

Only tests/, bench.py's cpu_baseline / reference arm and __graft_entry__.smoke() import this.
"""
from __future__ import annotations

import dataclasses
import types

import torch


@dataclasses.dataclass(frozen=True)
class Workload:
    """One hot-path configuration (names follow BASELINE.json `configs`)."""
    name: str
    B: int
    Ns: int
    Nt: int
    Ds: int
    Dt: int
    Lt: int
    H: int
    has_cls: bool
    student_depth: int = 12
    P: int = 4
    num_classes: int = 1000

    def token_layers(self) -> list[int]:
        # /root/reference/src/losses/combined.py:34-40 (Python banker's rounding)
        if self.P == 1:
            return [self.student_depth - 1]
        return [round(i * (self.student_depth - 1) / (self.P - 1)) for i in range(self.P)]


# BASELINE.json configs[0..4]
CONFIGS = {
    "cfg1": Workload("cfg1_vit_tiny_from_deit_small_b32", 32, 196, 196, 192, 384, 12, 6, True),
    "cfg2": Workload("cfg2_deit_tiny_from_deit_base_b256", 256, 196, 196, 192, 768, 12, 12, True),
    "cfg3": Workload("cfg3_resnet50_to_vit_small_b256", 256, 196, 49, 384, 2048, 1, 1, False),
    "cfg4": Workload("cfg4_vit_small_from_vit_large_b256", 256, 196, 196, 384, 1024, 24, 16, True),
    "cfg5": Workload("cfg5_deit_small_from_deit_base_384px_b128", 128, 576, 576, 384, 768, 12, 12, True),
}


def spiked(B, N, D, r, gen):
    basis = torch.linalg.qr(torch.randn(D, r, generator=gen))[0]
    x = (torch.randn(B, N, r, generator=gen) * 4.0 * torch.linspace(1.0, 0.2, r)) @ basis.T + torch.randn(B, N, D, generator=gen)
    return x.bfloat16()


def geometric(B, N, D, gen, rho=0.985, amp=3.0):
    basis = torch.linalg.qr(torch.randn(D, D, generator=gen))[0]
    x = (torch.randn(B, N, D, generator=gen) * amp * rho ** torch.arange(D)) @ basis.T
    return x.bfloat16()


def make_inputs(w: Workload, seed: int = 1234, batch: int | None = None, attn_dtype=torch.float32):
    """Returns dict(logits, targets, student{layer}, teacher{j}, attn{j}); tokens bf16, attention bf16-rounded
    values stored as `attn_dtype`.  Same generator call order as the survey probe that produced BASELINE.md's goldens."""
    B = batch if batch is not None else w.B
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, w.num_classes, generator=g)
    targets = torch.randint(0, w.num_classes, (B,), generator=g)
    student = {l: geometric(B, w.Ns, w.Ds, g) for l in w.token_layers()}
    teacher = {j: spiked(B, w.Nt, w.Dt, (16 + 4 * j) if w.Lt > 1 else 64, g) for j in range(w.Lt)}
    if w.has_cls:
        attn = {j: torch.softmax(2 * torch.randn(B, w.H, w.Nt + 1, w.Nt + 1, generator=g), -1).bfloat16().to(attn_dtype)
                for j in range(w.Lt)}
    else:
        attn = {j: (torch.ones(B, 1, w.Nt, w.Nt) / w.Nt).to(attn_dtype) for j in range(w.Lt)}
    return dict(logits=logits, targets=targets, student=student, teacher=teacher, attn=attn)


def module_config(w: Workload):
    return types.SimpleNamespace(num_extraction_points=w.P)


def standalone_inputs():
    """Seeded inputs of tests/golden/standalone.pt (the reference's standalone entry points; the fixture stores outputs only)."""
    w = Workload("standalone", 3, 48, 36, 40, 72, 3, 2, True, P=2)
    inp = make_inputs(w, seed=77)
    g = torch.Generator().manual_seed(78)
    teacher_aligned = spiked(w.B, w.Ns, w.Dt, 12, g)                       # teacher tokens already on the student's grid
    attn_same = torch.softmax(2 * torch.randn(w.B, w.H, w.Ns + 1, w.Ns + 1, generator=g), -1).bfloat16().float()
    attn_nocls = torch.softmax(2 * torch.randn(w.B, w.H, w.Nt, w.Nt, generator=g), -1).bfloat16().float()
    return w, inp, teacher_aligned, attn_same, attn_nocls
