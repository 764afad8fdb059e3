"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every declared symbol, host-only entry
points work without a GPU, and the Python module mirrors the reference's interface."""
import ctypes
import os
import re
import types

import pytest
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "basd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(basd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    names = declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/basd_b200.h but not exported"
    assert b"sm_100a" in lib.basd_version()


def test_workspace_bytes_is_host_only(lib):
    from vit_bias_aware_structural_distillation_b200._lib import Shape
    s = Shape(B=256, Ns=196, Nt=196, Ds=192, Dt=768, Lt=12, P=4, H=12, has_cls=1, act_dtype=1, attn_dtype=1, world_size=1)
    n = ctypes.c_size_t()
    assert lib.basd_workspace_bytes(ctypes.byref(s), ctypes.byref(n)) == 0
    assert 1e9 < n.value < 4e9          # sized for 180 GB of HBM, not for minimal footprint
    # BASELINE.json configs[2..4] are accepted: 7x7 CNN grid -> ViT-S, ViT-S <- ViT-L, 576 tokens at 384 px
    for kw in (dict(B=256, Ns=196, Nt=49, Ds=384, Dt=2048, Lt=1, H=1, has_cls=0), dict(B=256, Ns=196, Nt=196, Ds=384, Dt=1024, Lt=24, H=16, has_cls=1),
               dict(B=128, Ns=576, Nt=576, Ds=384, Dt=768, Lt=12, H=12, has_cls=1)):
        s = Shape(P=4, act_dtype=1, attn_dtype=1, world_size=1, **kw)
        assert lib.basd_workspace_bytes(ctypes.byref(s), ctypes.byref(n)) == 0, lib.basd_last_error().decode()
        assert n.value < 40e9
    # D_s > N_s - 1 with a finer teacher grid (ViT-S at 196 tokens <- a 384 px teacher): token-space form on the student's grid
    s = Shape(P=4, act_dtype=1, attn_dtype=1, world_size=1, B=64, Ns=196, Nt=576, Ds=384, Dt=768, Lt=12, H=12, has_cls=1)
    assert lib.basd_workspace_bytes(ctypes.byref(s), ctypes.byref(n)) == 0, lib.basd_last_error().decode()


@pytest.mark.parametrize("field,value,needle", [("Ds", 190, "multiples of 8"), ("Ds", 2048, "not supported"), ("Lt", 100, "Lt <="),
                                                ("world_size", 0, "world_size"), ("Nt", 256, "token-space Cholesky factor"), ("mode", 7, "mode")])
def test_unsupported_shapes_fail_loudly(lib, field, value, needle):
    from vit_bias_aware_structural_distillation_b200._lib import Shape
    kw = dict(B=8, Ns=196, Nt=196, Ds=192, Dt=768, Lt=12, P=4, H=12, has_cls=1, act_dtype=1, attn_dtype=1, world_size=1)
    kw[field] = value
    if field == "Nt":
        kw["Ds"], kw["Ns"] = 384, 300   # D_s > min(N_s, N_t) - 1 with more than 224 tokens on the coarser grid: the one remaining unbuilt size
    s = Shape(**kw)
    n = ctypes.c_size_t()
    assert lib.basd_workspace_bytes(ctypes.byref(s), ctypes.byref(n)) != 0
    assert needle in lib.basd_last_error().decode()


def test_module_mirrors_reference_interface(lib):
    import vit_bias_aware_structural_distillation_b200 as pkg
    torch.manual_seed(0)
    m = pkg.BASDLoss(nn.CrossEntropyLoss(), 192, 768, 12, 196, config=types.SimpleNamespace(num_extraction_points=4),
                     teacher_has_cls_token=True)
    assert m.token_layers == [0, 4, 7, 11]                                   # combined.py:34-40 (banker's rounding)
    sd = m.state_dict()
    assert set(sd) == {"layer_selector.log_temperatures", "layer_selector.proj_s", "layer_selector.proj_t"}
    assert sd["layer_selector.proj_s"].shape == (192, 192) and sd["layer_selector.proj_t"].shape == (192, 768)
    params = list(m.parameters())
    assert len(params) == 1 and params[0].shape == (4,)                       # trainer.py:74-76
    assert torch.allclose(m.layer_selector.temperatures, torch.ones(4))      # tau = 1.0
    pt = sd["layer_selector.proj_t"]
    assert torch.allclose(pt @ pt.T, torch.eye(192), atol=1e-5)
    assert m.layer_selector.subspace_ranks == {}
    m1 = pkg.BASDLoss(nn.CrossEntropyLoss(), 192, 768, 12, 196, config=types.SimpleNamespace(num_extraction_points=1),
                      teacher_has_cls_token=True)
    assert m1.token_layers == [11]


def test_same_seed_gives_reference_buffers(lib):
    """Built under the same torch.manual_seed the module reproduces the reference's proj_s / proj_t bit for bit
    (checked against the committed tiny golden, which stores nothing of them — so compare two constructions)."""
    import vit_bias_aware_structural_distillation_b200 as pkg
    cfg = types.SimpleNamespace(num_extraction_points=4)
    torch.manual_seed(0)
    a = pkg.BASDLoss(nn.CrossEntropyLoss(), 48, 96, 12, 64, config=cfg, teacher_has_cls_token=True)
    torch.manual_seed(0)
    ps = torch.empty(48, 48); pt = torch.empty(48, 96)
    nn.init.orthogonal_(ps); nn.init.orthogonal_(pt)
    assert torch.equal(a.layer_selector.proj_s, ps) and torch.equal(a.layer_selector.proj_t, pt)


def test_cpu_tensors_are_rejected_not_emulated(lib):
    import vit_bias_aware_structural_distillation_b200 as pkg
    from oracle import synth
    w = synth.Workload("t", 2, 64, 64, 48, 96, 3, 2, True)
    inp = synth.make_inputs(w)
    torch.manual_seed(0)
    m = pkg.BASDLoss(nn.CrossEntropyLoss(), w.Ds, w.Dt, 12, w.Ns, config=synth.module_config(w), teacher_has_cls_token=True)
    with pytest.raises(pkg.BasdError, match="CUDA device only"):
        m(inp["logits"], inp["targets"], inp["student"], inp["teacher"], inp["attn"])
    with pytest.raises(pkg.BasdError, match="CUDA tensor required"):
        pkg.marchenko_pastur_rank(torch.randn(100, 48))


def test_product_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "vit_bias_aware_structural_distillation_b200")
    for fn in os.listdir(pkg_dir):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg_dir, fn)).read()
            assert "oracle" not in src.replace("no autograd through LAPACK", ""), f"{fn} references oracle/"
