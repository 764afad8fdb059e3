"""Worker of tests/test_gpu_parity.py::test_two_rank_sharding_matches_single_process (launched with torchrun, 2 ranks).
Each rank takes its half of the cfg1 B=8 batch, runs the drop-in module (pooled statistics and d loss / d w are summed
across ranks inside it) and saves what it computed."""
import dataclasses
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import synth  # noqa: E402


def main():
    out_dir = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device(os.environ.get("BASD_SHARD_DEVICE", f"cuda:{os.environ.get('LOCAL_RANK', '0')}"))
    torch.cuda.set_device(dev)
    backend = "gloo" if "BASD_SHARD_DEVICE" in os.environ else "nccl"
    dist.init_process_group(backend)
    import vit_bias_aware_structural_distillation_b200 as pkg
    w = dataclasses.replace(synth.CONFIGS["cfg1"], B=8)
    inp = synth.make_inputs(w)
    per = w.B // world
    sl = slice(rank * per, (rank + 1) * per)
    torch.manual_seed(0)
    m = pkg.BASDLoss(nn.CrossEntropyLoss(label_smoothing=0.001), w.Ds, w.Dt, w.student_depth, w.Ns, config=synth.module_config(w),
                     teacher_has_cls_token=w.has_cls).to(dev)
    S = {l: v[sl].to(dev).requires_grad_() for l, v in inp["student"].items()}
    T = {j: v[sl].to(dev) for j, v in inp["teacher"].items()}
    A = {j: v[sl].to(dev) for j, v in inp["attn"].items()}
    logits = inp["logits"][sl].to(dev).requires_grad_()
    loss = m(logits, inp["targets"][sl].to(dev), S, T, A)
    loss.backward()
    torch.cuda.synchronize()
    # the global loss is the mean of the per-rank losses (equal shards)
    gl = loss.detach().clone()
    dist.all_reduce(gl)
    torch.save(dict(loss=(gl / world).cpu(), ranks=m.layer_selector.subspace_ranks, w=m.layer_selector.last_mixing_weights.cpu(),
                    grad_student={l: S[l].grad.float().cpu() for l in S},
                    grad_log_temperatures=m.layer_selector.log_temperatures.grad.cpu()), os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
