"""Development aid: phase timeline (clock64) of CTA 0 inside the four polar GEMM launches of Newton-Schulz step 3."""
import ctypes, os, sys
os.environ["BASD_POLAR_DBG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn as nn
import bench
import vit_bias_aware_structural_distillation_b200 as pkg
from oracle import synth
lib = pkg.load(); dev = torch.device("cuda:0")
w = bench.workload(256)
logits, targets, student, teacher, attn = bench.device_inputs(w, dev, 1)
torch.manual_seed(0)
m = pkg.BASDLoss(nn.CrossEntropyLoss(), w.Ds, w.Dt, w.student_depth, w.Ns, config=synth.module_config(w), teacher_has_cls_token=True).to(dev)
for _ in range(3):
    g = m.geo_loss(student, teacher, attn)
torch.cuda.synchronize()
names = ["G1 T=W Kt", "G2 A=T W^T", "G3 Bm=p(A)", "G4 W=Bm W"]
cols = ["prod_start", "prod_issued", "mma_wait_acc", "mma_start", "mma_issued", "epi_wait", "epi_got_acc", "epi_done"]
for which in range(4):
    buf = (ctypes.c_longlong * 128)()
    assert lib.basd_debug_polar_clocks(which, buf) == 0
    t = torch.tensor(list(buf)).view(16, 8)
    t0 = t[0, 0].item()
    print(f"== {names[which]}  (cycles since the producer started item 0; one row per item of CTA 0)")
    print("   " + " ".join(f"{c:>12s}" for c in cols))
    for i in range(14):
        print(f"{i:2d} " + " ".join(f"{(v - t0):12d}" for v in t[i].tolist()))
    if which == 1 and lib.basd_polar_launches_per_step(w.Ds, w.Ns) == 3:
        print("   (fused A/Bm kernel: columns = loads_start, tile1_stored, phase1_start, phase1_issued, copy_seen, phase2_issued, store_start, store_done)")
        continue
    if which == 2 and lib.basd_polar_launches_per_step(w.Ds, w.Ns) == 3:
        print("   (not launched: fused into the previous kernel)")
        continue
    e = t[15].tolist()
    print(f"   epilogue of item 5, column block 1 (cycles): tmem ld {e[1]-e[0]}, convert+stage {e[2]-e[1]}, fence+syncwarp {e[3]-e[2]}, "
          f"tma store issue {e[4]-e[3]}, wait_group.read {e[5]-e[4]}")
