"""Context number (BASELINE.md section 2 / SURVEY.md section 6): the UNMODIFIED reference BASDLoss (baseline/_ref/src/losses, eager
PyTorch -> cuBLAS / cuSOLVER) timed on the B200 itself, fp32 without autocast (the autocast path raises), on the
same synthetic inputs as bench.py.  Not the stated baseline (that is the CPU arm) - it shows what the stock code does on this GPU.
usage (under gpurun): python tools/ref_gpu_probe.py [--workload cfg2] [--batch 256] [--steps 3]"""
import argparse
import dataclasses
import json
import os
import sys
import time

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
from oracle import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    from src.losses.combined import BASDLoss as RefLoss
    w = synth.CONFIGS[args.workload]
    if args.batch:
        w = dataclasses.replace(w, B=args.batch)
    dev = torch.device("cuda:0")
    inp = synth.make_inputs(w)
    torch.manual_seed(0)
    m = RefLoss(nn.CrossEntropyLoss(label_smoothing=0.001), w.Ds, w.Dt, w.student_depth, w.Ns, config=synth.module_config(w),
                teacher_has_cls_token=w.has_cls).to(dev)
    S = {l: v.float().to(dev).requires_grad_() for l, v in inp["student"].items()}
    T = {j: v.float().to(dev) for j, v in inp["teacher"].items()}
    A = {j: v.float().to(dev) for j, v in inp["attn"].items()}
    logits = inp["logits"].to(dev).requires_grad_()
    targets = inp["targets"].to(dev)

    def step():
        for t in S.values():
            t.grad = None
        m.zero_grad(set_to_none=True)
        loss = m(logits, targets, S, T, A)
        loss.backward()
        return loss

    loss = step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        loss = step()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    sec = sum(ts) / len(ts)
    print(json.dumps({"impl": "reference on the B200 (eager PyTorch, cuBLAS + cuSOLVER, fp32)", "workload": w.name, "batch": w.B,
                      "ms_per_step": sec * 1e3, "samples_per_s": w.B / sec, "loss": float(loss), "steps": args.steps,
                      "peak_memory_gb": torch.cuda.max_memory_allocated() / 1e9, "torch": torch.__version__}), flush=True)


if __name__ == "__main__":
    main()
