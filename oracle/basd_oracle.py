"""CPU restatement of the BASD distillation-loss hot path (TEST INFRASTRUCTURE — the checker, never the product).

What it restates (file:line under /root/reference/src/losses/):
  * marchenko_pastur_rank                layer_selector.py:8-20
  * _grassmann_subspace                  layer_selector.py:23-37
  * GrassmannianLayerSelector.forward    layer_selector.py:116-152 (+ _estimate_ranks :69-74,
                                         _mix_for_student_layer :76-114)
  * _align_token_count                   combined.py:9-14
  * geometric_relational_loss            relational.py:5-50
  * BASDLoss.forward (CE + UW-SO)        combined.py:48-85

The arithmetic that lives in a third-party dependency (torch.linalg.{eigvalsh,svd,svdvals,matrix_norm}
-> LAPACK via MKL; the reference pins torch==2.10.0 in pyproject.toml:6, this image has 2.11.0) is
called here through the same torch entry points, in fp32 (the reference's precision) or fp64 (the
conditioning referee of SURVEY.md §8c).

PINNING: the reference has no tests, golden vectors or fixtures for this path, so the pin
is the reference itself executed in the build container: tests/golden/*.pt are produced by
oracle/make_golden.py from the UNMODIFIED /root/reference module, and tests/test_oracle_pinned.py checks
this restatement against them (loss, ranks, mixing weights, temperature gradients, student-gradient
projections).  The restatement differs from the reference only by reorderings that are exact in real
arithmetic: the attention importance row is reduced per teacher layer BEFORE the layer mixing
 instead of mixing full [B,H,N+1,N+1] maps.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def mp_rank(features: torch.Tensor) -> int:
    """layer_selector.py:8-20 — uncentred covariance, LOWER median (torch.median), strict '>' count."""
    M, D = features.shape
    q = D / M
    cov = features.T @ features / M if M >= D else features @ features.T / M
    ev = torch.linalg.eigvalsh(cov)
    sigma2 = ev.median().item()
    lam_plus = sigma2 * (1 + q ** 0.5) ** 2
    return int((ev > lam_plus).sum().item())


def interp_linear_1d(x: torch.Tensor, n_out: int) -> torch.Tensor:
    """combined.py:9-14 / relational.py:29-32 — 2-tap linear resampling along dim 1 of [B, N, ...],
    align_corners=False.  Written out explicitly (no F.interpolate) so the CUDA
    kernel's index arithmetic has a plain statement to be compared with."""
    n_in = x.shape[1]
    if n_in == n_out:
        return x
    pos = ((torch.arange(n_out, dtype=torch.float64) + 0.5) * (n_in / n_out) - 0.5).clamp(min=0.0)
    i0 = pos.floor().long().clamp(max=n_in - 1)
    i1 = (i0 + 1).clamp(max=n_in - 1)
    lam = (pos - i0.double()).to(x.dtype)
    shape = [1, n_out] + [1] * (x.dim() - 2)
    lam = lam.view(shape)
    return (1 - lam) * x[:, i0] + lam * x[:, i1]


def importance_rows(attn: torch.Tensor, has_cls: bool) -> torch.Tensor:
    """relational.py:22-27 — CLS row averaged over heads, or mean over (heads, queries)."""
    if has_cls:
        return attn[:, :, 0, 1:].mean(dim=1)
    return attn.mean(dim=(1, 2))


def geometric_relational_loss(student_tokens, teacher_tokens, teacher_attn, *, has_cls, dtype=torch.float32):
    """relational.py:5-50 on its own (one student / teacher pair, teacher tokens already on the student's token grid)."""
    s, t = student_tokens.to(dtype), teacher_tokens.to(dtype)
    imp = importance_rows(teacher_attn.to(dtype), has_cls)               # relational.py:22-27
    imp = interp_linear_1d(imp, s.shape[1])                              # relational.py:29-32
    a = imp / imp.sum(dim=-1, keepdim=True)                              # relational.py:34
    mu_s = (a.unsqueeze(-1) * s).sum(dim=1, keepdim=True)                # relational.py:36-39
    mu_t = (a.unsqueeze(-1) * t).sum(dim=1, keepdim=True)
    rt = a.unsqueeze(-1).sqrt()                                          # relational.py:41-43
    s_w, t_w = rt * (s - mu_s), rt * (t - mu_t)
    tr_s, tr_t = (s_w * s_w).sum(dim=(1, 2)), (t_w * t_w).sum(dim=(1, 2))     # relational.py:45-46
    nuc = torch.linalg.matrix_norm(torch.bmm(s_w.transpose(1, 2), t_w), ord="nuc")   # relational.py:47-48
    return (tr_s + tr_t - 2.0 * nuc).mean()                              # relational.py:50


def selector_forward(student, teacher, attn, proj_s, proj_t, log_temperatures, extraction_indices, dtype=torch.float32):
    """GrassmannianLayerSelector.forward (layer_selector.py:116-152) on its own: (mixed_teachers, mixed_attentions, ranks)."""
    t_idx = sorted(teacher.keys())
    Ds = proj_s.shape[0]
    Pt, Ps = proj_t.to(dtype), proj_s.to(dtype)
    ranks, bases, svals = {}, {}, {}
    with torch.no_grad():
        for j in t_idx:                                                  # ls:69-74, 131-138
            z = teacher[j].to(dtype).reshape(-1, teacher[j].shape[2]) @ Pt.T
            ranks[j] = min(mp_rank(z), Ds - 1)
            zc = z - z.mean(dim=0, keepdim=True)
            _, S, Vt = torch.linalg.svd(zc, full_matrices=False)
            bases[j], svals[j] = Vt[: ranks[j]].T, S[: ranks[j]]
    T = torch.stack([teacher[j].to(dtype) for j in t_idx])              # ls:128-129
    A = torch.stack([attn[j].to(dtype) for j in t_idx])
    mixed_t, mixed_a = {}, {}
    for i, layer in enumerate(extraction_indices):                       # ls:76-114
        zs = student[layer].to(dtype).reshape(-1, Ds) @ Ps.T
        zs = zs - zs.mean(dim=0, keepdim=True)
        _, _, Vts = torch.linalg.svd(zs, full_matrices=False)
        d2 = []
        for j in t_idx:
            sig = torch.linalg.svdvals(Vts[: ranks[j]] @ bases[j])
            theta = torch.acos(sig.clamp(max=1.0 - torch.finfo(sig.dtype).eps))
            d2.append((svals[j] * theta.pow(2)).sum() / svals[j].sum())
        w = F.softmax(-torch.stack(d2) / F.softplus(log_temperatures.to(dtype)[i]), dim=0).to(T.dtype)
        mixed_t[layer] = (w.view(-1, 1, 1, 1) * T).sum(dim=0)
        mixed_a[layer] = (w.view(-1, 1, 1, 1, 1) * A).sum(dim=0)
    return mixed_t, mixed_a, ranks


def forward(student, teacher, attn, proj_s, proj_t, log_temperatures, token_layers, *, has_cls,
            n_student_tokens, dtype=torch.float32, ce_loss=None, detach_weights=False):
    """Whole hot path.  `student`/`teacher`/`attn` are dicts like the reference takes.
    Returns a dict of invariants; `geo` (and `loss` if ce_loss given) carry autograd graphs back to
    the student tensors passed in and to `log_temperatures`."""
    dt = dtype
    t_idx = sorted(teacher.keys())
    Ds = proj_s.shape[0]
    Pt = proj_t.to(dt)
    Ps = proj_s.to(dt)
    ranks, bases, svals = {}, {}, {}
    with torch.no_grad():
        for j in t_idx:                                             # ls:69-74 and ls:131-138
            z = teacher[j].to(dt).reshape(-1, teacher[j].shape[2]) @ Pt.T
            ranks[j] = min(mp_rank(z), Ds - 1)
            zc = z - z.mean(dim=0, keepdim=True)
            _, S, Vt = torch.linalg.svd(zc, full_matrices=False)
            bases[j] = Vt[: ranks[j]].T
            svals[j] = S[: ranks[j]]
        T = torch.stack([teacher[j].to(dt) for j in t_idx])          # [Lt,B,Nt,Dt]
        rows = torch.stack([importance_rows(attn[j].to(dt), has_cls) for j in t_idx])   # [Lt,B,Nt]

    out = dict(ranks=ranks, d2=[], cos=[], w=[], nuc=[], tr_s=[], tr_t=[], geo_i=[])
    for i, layer in enumerate(token_layers):
        s_in = student[layer]
        s = s_in.to(dt)
        zs = s.reshape(-1, Ds) @ Ps.T                                # ls:86-92
        zs = zs - zs.mean(dim=0, keepdim=True)
        _, _, Vts = torch.linalg.svd(zs, full_matrices=False)
        d2, cosines = [], []
        for j in t_idx:                                              # ls:95-105
            k = ranks[j]
            sig = torch.linalg.svdvals(Vts[:k] @ bases[j])
            theta = torch.acos(sig.clamp(max=1.0 - torch.finfo(sig.dtype).eps))
            d2.append((svals[j] * theta.pow(2)).sum() / svals[j].sum())
            cosines.append(sig.detach())
        d2 = torch.stack(d2)
        tau = F.softplus(log_temperatures.to(dt)[i])                 # ls:65-67,107
        w = F.softmax(-d2 / tau, dim=0)                              # ls:108
        out["d2"].append(d2.detach()); out["cos"].append(cosines); out["w"].append(w.detach())
        if detach_weights:
            w = w.detach()
        wt = w.to(T.dtype)                                           # ls:110
        mixed = (wt.view(-1, 1, 1, 1) * T).sum(dim=0)                # ls:111
        imp = (wt.view(-1, 1, 1) * rows).sum(dim=0)                  # ls:112 + relational.py:22-27 (commute)
        mixed = interp_linear_1d(mixed, n_student_tokens)            # combined.py:63-67
        imp = interp_linear_1d(imp, s.shape[1])                      # relational.py:29-32
        a = imp / imp.sum(dim=-1, keepdim=True)                      # relational.py:34
        mu_s = (a.unsqueeze(-1) * s).sum(dim=1, keepdim=True)        # relational.py:36-39
        mu_t = (a.unsqueeze(-1) * mixed).sum(dim=1, keepdim=True)
        s_c, t_c = s - mu_s, mixed - mu_t
        rt = a.unsqueeze(-1).sqrt()                                  # relational.py:41-43
        s_w, t_w = rt * s_c, rt * t_c
        tr_s = (s_w * s_w).sum(dim=(1, 2))                           # relational.py:45-46
        tr_t = (t_w * t_w).sum(dim=(1, 2))
        cross = torch.bmm(s_w.transpose(1, 2), t_w)                  # relational.py:47
        nuc = torch.linalg.matrix_norm(cross, ord="nuc")             # relational.py:48
        out["nuc"].append(nuc.detach()); out["tr_s"].append(tr_s.detach()); out["tr_t"].append(tr_t.detach())
        out["geo_i"].append((tr_s + tr_t - 2.0 * nuc).mean())        # relational.py:50
    geo = torch.stack(out["geo_i"]).mean()                           # combined.py:76
    out["geo"] = geo
    out["geo_i"] = [g.detach() for g in out["geo_i"]]
    out["w"] = torch.stack(out["w"]); out["d2"] = torch.stack(out["d2"])
    out["nuc"] = torch.stack(out["nuc"]); out["tr_s"] = torch.stack(out["tr_s"]); out["tr_t"] = torch.stack(out["tr_t"])
    if ce_loss is not None:
        out["ce"] = ce_loss
        out["loss"] = uwso([ce_loss.to(dt), geo])
    return out


def uwso(vals):
    """combined.py:78-85 — inverse-loss weights on detached values."""
    eps = torch.finfo(vals[0].dtype).eps
    inv = torch.stack([1.0 / v.detach().clamp(min=eps) for v in vals])
    w = inv / inv.sum()
    return sum(w[i] * vals[i] for i in range(len(vals)))


def run_case(inputs, proj_s, proj_t, log_temperatures, token_layers, *, has_cls, n_student_tokens,
             dtype=torch.float32, label_smoothing=0.0, detach_weights=False):
    """Forward + autograd backward on fresh leaf copies; returns invariants plus gradients."""
    student = {l: v.detach().to(dtype).clone().requires_grad_() for l, v in inputs["student"].items()}
    teacher = {j: v.detach().to(dtype) for j, v in inputs["teacher"].items()}
    attn = {j: v.detach().to(dtype) for j, v in inputs["attn"].items()}
    logt = log_temperatures.detach().to(dtype).clone().requires_grad_()
    logits = inputs["logits"].detach().to(dtype).clone().requires_grad_()
    ce = F.cross_entropy(logits, inputs["targets"], label_smoothing=label_smoothing)
    out = forward(student, teacher, attn, proj_s, proj_t, logt, token_layers, has_cls=has_cls,
                  n_student_tokens=n_student_tokens, dtype=dtype, ce_loss=ce, detach_weights=detach_weights)
    out["loss"].backward()
    out["grad_student"] = {l: v.grad for l, v in student.items()}
    out["grad_log_temperatures"] = logt.grad if logt.grad is not None else torch.zeros_like(logt)
    out["grad_logits"] = logits.grad
    out["loss"] = out["loss"].detach(); out["geo"] = out["geo"].detach(); out["ce"] = out["ce"].detach()
    return out
