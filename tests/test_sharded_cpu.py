"""world_size-2 `gloo` tests (CPU) of the batch-sharded path.

The product path has no CPU implementation, so what runs here is oracle/kernel_model.py — the stage-by-stage statement
of the algorithm the CUDA kernels implement — with REAL torch.distributed all-reduces placed exactly where
vit_bias_aware_structural_distillation_b200/loss.py places them (pooled statistics after phase 1, d loss / d w after
phase 3, the two UW-SO scalars).  It pins the collective design: what is summed, how M and the gradients are scaled,
and that two ranks on half batches reproduce ONE process of the reference restatement on the concatenated batch.
The same comparison on the real kernels is tests/test_gpu_parity.py::test_two_rank_sharding_matches_single_process."""
import dataclasses
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from oracle import basd_oracle as O
from oracle import kernel_model as K
from oracle import synth

W = synth.Workload("shard_cpu", 6, 40, 40, 24, 48, 3, 2, True)


def _buffers(w):
    torch.manual_seed(0)
    ps = torch.empty(w.Ds, w.Ds); pt = torch.empty(w.Ds, w.Dt)
    nn.init.orthogonal_(ps); nn.init.orthogonal_(pt)
    return ps, pt, torch.full((w.P,), math.log(math.exp(1.0) - 1))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    inp = synth.make_inputs(W)
    per = W.B // world
    sl = slice(rank * per, (rank + 1) * per)
    shard = dict(logits=inp["logits"][sl], targets=inp["targets"][sl], student={l: v[sl] for l, v in inp["student"].items()},
                 teacher={j: v[sl] for j, v in inp["teacher"].items()}, attn={j: v[sl] for j, v in inp["attn"].items()})
    ps, pt, logt = _buffers(W)
    ce = torch.nn.functional.cross_entropy(shard["logits"].double(), shard["targets"], label_smoothing=0.001)
    calls = []

    def allreduce(t):
        calls.append(t.numel())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    out = K.forward_backward(shard, ps, pt, logt, W.token_layers(), has_cls=W.has_cls, n_student_tokens=W.Ns, dtype=torch.float64,
                             ce=ce, allreduce=allreduce, world=world)
    gl = out["loss"].detach().clone()
    dist.all_reduce(gl)
    torch.save(dict(loss=gl / world, ranks=out["ranks"], w=out["w"], grad_student=out["grad_student"],
                    grad_log_temperatures=out["grad_log_temperatures"], n_collectives=len(calls), payload=sum(calls)),
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gloo_ranks_match_single_process_reference(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    parts = [torch.load(tmp_path / f"rank{r}.pt", weights_only=False) for r in range(world)]
    inp = synth.make_inputs(W)
    ps, pt, logt = _buffers(W)
    ref = O.run_case(inp, ps, pt, logt, W.token_layers(), has_cls=W.has_cls, n_student_tokens=W.Ns, dtype=torch.float64,
                     label_smoothing=0.001)
    per = W.B // world
    for r, p in enumerate(parts):
        assert p["ranks"] == ref["ranks"]                                          # pooled over BOTH shards
        assert (p["w"] - ref["w"]).abs().max() < 1e-6         # acos near 1 amplifies the fp64 Gram-vs-SVD rounding
        # UW-SO on global means: the mean of the per-rank losses is the single-process loss
        assert abs(p["loss"].item() - ref["loss"].item()) < 1e-8 * abs(ref["loss"].item())
        assert ((p["grad_log_temperatures"] - ref["grad_log_temperatures"]).abs() / ref["grad_log_temperatures"].abs()).max() < 1e-4
        for l in W.token_layers():
            mine = p["grad_student"][l] / world                                    # rank-local mean -> global mean
            want = ref["grad_student"][l][r * per:(r + 1) * per]
            assert ((mine - want).norm() / want.norm()).item() < 1e-5
    # every rank issued the same collectives: (G, c) per teacher layer and student point, gw per point, one UW-SO pair
    expect = 2 * (W.Lt + W.P) + W.P + 1
    assert parts[0]["n_collectives"] == parts[1]["n_collectives"] == expect
    assert parts[0]["payload"] == (W.Lt + W.P) * (W.Ds * W.Ds + W.Ds) + W.P * W.Lt + 2       # == the "stats" + "gw" views + 2 scalars


def test_sharding_is_exact_only_with_pooled_statistics(tmp_path):
    """Negative control: without the statistics all-reduce (each rank pooling over its own shard, which is what running
    the unmodified reference under DDP would do) the mixing weights differ from the
    single-process result - the collective is not optional."""
    inp = synth.make_inputs(W)
    ps, pt, logt = _buffers(W)
    ref = O.run_case(inp, ps, pt, logt, W.token_layers(), has_cls=W.has_cls, n_student_tokens=W.Ns, dtype=torch.float64, label_smoothing=0.001)
    half = dict(logits=inp["logits"][:3], targets=inp["targets"][:3], student={l: v[:3] for l, v in inp["student"].items()},
                teacher={j: v[:3] for j, v in inp["teacher"].items()}, attn={j: v[:3] for j, v in inp["attn"].items()})
    alone = K.forward_backward(half, ps, pt, logt, W.token_layers(), has_cls=W.has_cls, n_student_tokens=W.Ns, dtype=torch.float64)
    assert (alone["w"] - ref["w"]).abs().max() > 1e-4
