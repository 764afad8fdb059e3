"""Development aid: which torch ops run on the GPU inside one step (besides the library's own kernels)?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn as nn
import bench
import vit_bias_aware_structural_distillation_b200 as pkg
from oracle import synth
dev = torch.device("cuda:0")
w = bench.workload(256)
logits, targets, student, teacher, attn = bench.device_inputs(w, dev, 1)
torch.manual_seed(0)
m = pkg.BASDLoss(nn.CrossEntropyLoss(label_smoothing=0.001), w.Ds, w.Dt, w.student_depth, w.Ns, config=synth.module_config(w), teacher_has_cls_token=True).to(dev)
logits.requires_grad_()
for t in student.values(): t.requires_grad_()
def step():
    for t in student.values(): t.grad = None
    m.zero_grad(set_to_none=True); logits.grad = None
    loss = m(logits, targets, student, teacher, attn); loss.backward(); return loss
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages(group_by_stack_n=6).table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
