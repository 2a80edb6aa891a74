"""CPU-only checks of the drop-in boundary: config loading, registries, module/state_dict layout, and that the C-ABI
library loads and exports every symbol include/ucod_b200.h declares (no compute calls — there is no GPU here)."""
import ctypes
import json
import re
from pathlib import Path
from types import SimpleNamespace

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
GOLD = ROOT / "tests" / "golden"


def _plain(d):
    return {k: (_plain(v) if isinstance(v, dict) else (list(v) if isinstance(v, tuple) else v)) for k, v in d.items()}


# ---- engine.config.CfgNode (reference engine/config/config.py:66-191,244-263) ------------------------------
@pytest.mark.parametrize("name", ["UCOD-DPL_dinov1.py", "UCOD-DPL_dinov2.py", "CORAL_dinov1.py", "CORAL_dinov2.py"])
def test_configs_resolve_like_the_reference(name):
    from ucod_dpl_b200.engine.config.config import CfgNode
    gold = json.loads((GOLD / "configs.json").read_text())[name]
    cfg = CfgNode.load_with_base(str(ROOT / "configs" / "uscod" / name))
    assert _plain(cfg) == gold
    assert cfg.model_cfg.dim == 768 and cfg["model_cfg"]["dim"] == 768
    assert isinstance(cfg.dataset_cfg.valset_cfg.image_size, tuple)
    ref = Path("/root/reference/configs/uscod") / name  # the reference's own files load unchanged (build box only)
    if ref.exists():
        assert _plain(CfgNode.load_with_base(str(ref))) == gold


def test_cfgnode_semantics(tmp_path):
    from ucod_dpl_b200.engine.config.config import CfgNode
    c = CfgNode({"a": {"b": 1, "c": [1, 2]}, "d": "x"})
    assert c.a.b == 1 and c.get("zzz", 5) == 5
    c.merge_from_list(["a.b", "3", "d", "y"])
    assert c.a.b == 3 and c.d == "y"
    with pytest.raises(KeyError):
        c.merge_from_list(["a.q.z", "1"])
    c2 = c.clone()
    c2.a.b = 9
    assert c.a.b == 3
    c.freeze()
    with pytest.raises(AttributeError):
        c.a.b = 4
    c.defrost()
    c.a.b = 4
    dumped = c.dump()
    assert "a:" in dumped and "b: 4" in dumped
    base = tmp_path / "base.py"
    base.write_text("cfg = dict(x=1, n=dict(p=1, q=2))\n")
    child = tmp_path / "child.py"
    child.write_text("cfg = dict(_BASE_='base.py', n=dict(q=5), y=2)\n")
    m = CfgNode.load_with_base(str(child))
    assert _plain(m) == {"x": 1, "n": {"p": 1, "q": 5}, "y": 2}


# ---- engine.registry (reference engine/registry/registry.py:36-92, root.py:3-6) ----------------------------
def test_registries_are_populated():
    import ucod_dpl_b200.dropin  # noqa: F401  (imports every hot-path module so that it registers itself)
    from ucod_dpl_b200.engine.registry import BACKBONE_REGISTRY, DATASET_REGISTRY, HOOK_REGISTRY, MODULE_REGISTRY, Registry
    for n in ("baseline", "RevDecoder", "Discriminator", "SparseRefiner"):
        assert n in MODULE_REGISTRY
    assert "backbone" in BACKBONE_REGISTRY
    assert len(DATASET_REGISTRY) == 0 and len(HOOK_REGISTRY) == 0
    r = Registry("t")

    @r.register()
    class A:
        pass

    def f():
        return 1
    r.register(f)
    assert r.get("A") is A and r.get("f") is f and dict(iter(r)) == {"A": A, "f": f}
    with pytest.raises(KeyError):
        r.get("nope")
    with pytest.raises(AssertionError):
        r.register(f)


# ---- model/state_dict layout --------------------------------------------------------------
@pytest.mark.parametrize("kind", ["dinov1", "dinov2"])
def test_shipped_checkpoints_load_strict(kind):
    from safetensors.torch import load_file
    from ucod_dpl_b200.models.uscod import baseline
    sd = load_file(str(ROOT / "weights" / f"UCOD_DPL_{kind}.safetensors"))
    m = baseline(SimpleNamespace(dim=768))
    missing, unexpected = m.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    assert sorted(m.state_dict()) == sorted(sd)
    assert sum(v.numel() for v in sd.values()) == 197380


def test_discriminator_layout():
    from oracle import decoder as odec
    from ucod_dpl_b200.models.discriminator import Discriminator
    d = Discriminator(SimpleNamespace(dis_use_features=False, dim=768, feature_size=68))
    assert all(not p.requires_grad for p in d.parameters())
    d.load_state_dict(odec.random_discriminator_state_dict(68, seed=1), strict=True)
    assert d.linear.in_features == 2312


def test_refiner_state_dict_layout():
    import numpy as np
    from ucod_dpl_b200.models.UDLR import SparseRefiner
    from ucod_dpl_b200.synth import random_refiner_state_dict
    r = SparseRefiner.from_config(SimpleNamespace(window_size=3, threshold=0.0015))
    want = list(np.load(GOLD / "coral.npz")["state_dict_keys"])   # keys of the reference module
    assert sorted(r.state_dict().keys()) == want
    r.load_state_dict(random_refiner_state_dict(0), strict=True)


def test_reference_module_paths_alias():
    """`ucod_dpl_b200.dropin.install()` makes the reference's import paths resolve to this package."""
    import sys
    import ucod_dpl_b200.dropin as dropin
    saved = {k: sys.modules.get(k) for k in dropin.ALIASES}
    try:
        dropin.install()
        from models.uscod import baseline  # noqa
        from models.discriminator import Discriminator  # noqa
        from models.UDLR import SparseRefiner  # noqa
        from models.modules.CSF import CSF  # noqa
        from data.utils.feature_extractor import backbone  # noqa
        from data.utils.found_bkg_mask import compute_img_bkg_seg  # noqa
        from engine.config import CfgNode  # noqa
        from engine.registry import MODULE_REGISTRY  # noqa
        import ucod_dpl_b200.models.uscod as ours
        assert baseline is ours.baseline
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


# ---- the C-ABI library ------------------------------------------------------------------------------------
def _declared_symbols():
    text = (ROOT / "include" / "ucod_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ucod_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from ucod_dpl_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ucod_b200.h but not exported"
    lib.ucod_abi_version.restype = ctypes.c_int
    m = re.search(r"#define UCOD_B200_ABI_VERSION (\d+)", (ROOT / "include" / "ucod_b200.h").read_text())
    assert lib.ucod_abi_version() == int(m.group(1))


def test_no_cpu_fallback():
    from ucod_dpl_b200 import ops
    from ucod_dpl_b200._lib import UcodError
    with pytest.raises(UcodError):
        ops.upsample_bilinear(torch.zeros(1, 4, 4), (8, 8))
    with pytest.raises(UcodError):
        ops.refine_small_components(torch.zeros(1, 16, 16, dtype=torch.uint8))
    with pytest.raises(UcodError):
        ops.pseudo_label_score(torch.zeros(1, 12, 256), torch.zeros(1, 256, 768), 0.6)


def test_argument_errors_do_not_need_a_gpu():
    """Validation failures come back as status != 0 with a message (no exceptions across the ABI)."""
    from ucod_dpl_b200 import _lib
    lib = _lib.load()
    lib.ucod_gemm_bf16.restype = ctypes.c_int
    rc = lib.ucod_gemm_bf16(None, 8, None, 8, 0, 128, 64, 0, None, None, 128, None)
    assert rc != 0
    assert b"gemm" in lib.ucod_last_error()


def test_cache_format_interop_with_reference(tmp_path):
    """MetaListPickleIO writes what the reference reads and reads what the reference writes
    (engine/utils/fileio/backend/ioctl/pickleio.py:54-142)."""
    from ucod_dpl_b200.engine.utils.fileio import MetaListPickleIO
    items = [torch.full((1, 16, 16), float(i)) for i in range(5)]
    io = MetaListPickleIO(base_path=tmp_path / "ours")
    assert io.mode == "w"
    io.dump_list(items)
    assert sorted(p.name for p in (tmp_path / "ours").iterdir()) == ["data_%d.pkl" % i for i in range(5)] + ["index.json"]
    assert json.loads((tmp_path / "ours" / "index.json").read_text()) == {str(i): f"data_{i}.pkl" for i in range(5)}
    rd = MetaListPickleIO(base_path=tmp_path / "ours")
    assert rd.mode == "r" and rd.len() == 5 and torch.equal(rd.read_file(3), items[3])
    with pytest.raises(AssertionError):
        rd.write_file(9, items[0])
    (tmp_path / "ours" / "data_2.pkl").unlink()                      # a missing item invalidates the cache
    assert MetaListPickleIO(base_path=tmp_path / "ours").mode == "w"
    if not Path("/root/reference/engine/utils/fileio/backend/ioctl/pickleio.py").exists():
        return                                                       # GPU box: no reference tree
    # build box: drive the REAL reference class in a subprocess (its package needs the shims of tools/make_golden.py)
    import subprocess, sys
    io2 = MetaListPickleIO(base_path=tmp_path / "ours2")
    io2.dump_list(items)
    code = f"""
import sys, torch
sys.path.insert(0, {str(ROOT)!r})
import tools.make_golden as mg
mg.install_shims()
from engine.utils.fileio.backend.ioctl.pickleio import MetaListPickleIO as Ref
rd = Ref(base_path={str(tmp_path / 'ours2')!r})
assert rd.mode == 'r' and rd.len() == 5 and float(rd.read_file(4)[0, 0, 0]) == 4.0
wr = Ref(base_path={str(tmp_path / 'theirs')!r})
wr.dump_list([torch.full((1, 16, 16), float(i)) for i in range(5)])
print('REF_OK')
"""
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "REF_OK" in out.stdout, out.stderr[-2000:]
    ours_rd = MetaListPickleIO(base_path=tmp_path / "theirs")
    assert ours_rd.mode == "r" and ours_rd.len() == 5 and torch.equal(ours_rd.read_file(1), items[1])


def test_backbone_checkpoint_discovery_and_random_opt_in(tmp_path, monkeypatch):
    """`load_vit_state_dict`: picks the checkpoint that matches the requested architecture (DINO ViT-B/8 vs DINOv2-B/14
    both live under ./weights for the dinov1 configs), and refuses to fall back to random weights silently."""
    from types import SimpleNamespace

    import pytest
    from safetensors.torch import save_file

    from ucod_dpl_b200.data.utils import feature_extractor as fe
    from ucod_dpl_b200.synth import random_vit_state_dict
    from ucod_dpl_b200.vit import spec_for

    def tiny(kind):  # only the tensors the matcher looks at
        sd = random_vit_state_dict(spec_for(kind), seed=1)
        keep = ("embeddings.patch_embeddings.projection.weight", "embeddings.cls_token",
                "encoder.layer.0.attention.attention.key.weight", "encoder.layer.0.layer_scale1.lambda1")
        return {k: v.contiguous() for k, v in sd.items() if k in keep}

    (tmp_path / "a_dinov2").mkdir(), (tmp_path / "b_dino").mkdir()
    save_file(tiny("dinov2"), str(tmp_path / "a_dinov2" / "model.safetensors"))
    save_file(tiny("dinov1"), str(tmp_path / "b_dino" / "model.safetensors"))
    for kind, name, patch in (("dinov2", "facebook/dinov2-base", 14), ("dinov1", "facebook/dino-vitb8", 8)):
        cfg = SimpleNamespace(type=kind, backbone=name, backbone_weights=None, backbone_weight_base=str(tmp_path))
        sd = fe.load_vit_state_dict(cfg)
        assert sd["embeddings.patch_embeddings.projection.weight"].shape[-1] == patch
    empty = SimpleNamespace(type="dinov2", backbone="facebook/dinov2-base", backbone_weights=None,
                            backbone_weight_base=str(tmp_path / "nothing"))
    monkeypatch.delenv(fe.ALLOW_RANDOM_ENV, raising=False)
    with pytest.raises(FileNotFoundError):
        fe.load_vit_state_dict(empty)
    assert "embeddings.cls_token" in fe.load_vit_state_dict(empty, allow_random_init=True)
    monkeypatch.setenv(fe.ALLOW_RANDOM_ENV, "1")
    assert "embeddings.cls_token" in fe.load_vit_state_dict(empty)
