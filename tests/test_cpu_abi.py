"""C-ABI library: loads on a CPU box, exports every symbol include/detikzify_b200.h declares, and the
layout helpers (no GPU compute) agree with the model shapes. The product path must fail loudly without CUDA."""
import ctypes as C
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    from detikzify_b200 import _lib
    lib = _lib.load_library()
    header = (ROOT / "include" / "detikzify_b200.h").read_text()
    declared = set(re.findall(r"DTK_API\s+[\w\s\*]+?\b(dtk_\w+)\s*\(", header))
    assert len(declared) >= 24
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert lib.dtk_abi_version() == 2


@pytest.mark.parametrize("name,arena_gb", [("nllg/detikzify-ds-1.3b", 3.5637), ("nllg/detikzify-ds-7b", 14.3659)])
def test_weight_table_and_decode_bytes(name, arena_gb):
    from detikzify_b200 import _lib
    from detikzify_b200.engine import to_c_config, weight_table
    from detikzify_b200.model.configuration import preset
    from detikzify_b200.model.weights import param_count
    lib = _lib.load_library()
    cfg = preset(name)
    cc = to_c_config(cfg)
    table = weight_table(cc)
    names = [t.name.decode() for t in table]
    assert len(set(names)) == len(names)
    # table covers exactly the canonical parameters (+ K padding of the patch-embed weight: 588 -> 640 columns)
    n_elems = sum(t.rows * t.cols for t in table)
    assert n_elems == param_count(cfg) + cfg.vision_config.hidden_size * (640 - 588)
    for a, b in zip(table, table[1:]):
        assert a.offset % 256 == 0 and a.offset + a.nbytes <= b.offset
    assert abs(lib.dtk_arena_bytes(C.byref(cc)) / 1e9 - arena_gb) < 1e-3
    # algorithmic bytes of one decoded token: weights once + KV rows read
    H, I, L, V = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers, cfg.vocab_size
    w = 2 * (L * (4 * H * H + 3 * H * I) + V * H)
    for T in (243, 2048):
        assert lib.dtk_decode_bytes(C.byref(cc), T) == w + T * 2 * L * H * 2


def test_arena_packing_round_trip():
    from detikzify_b200.engine import pack_arena, to_c_config, weight_table
    from detikzify_b200.model.configuration import preset
    from detikzify_b200.model.weights import random_init
    cfg = preset("tiny")
    sd = random_init(cfg, seed=3)
    arena = pack_arena(cfg, sd)
    info = {t.name.decode(): t for t in weight_table(to_c_config(cfg))}

    def view(n):
        t = info[n]
        return arena[t.offset // 2: t.offset // 2 + t.rows * t.cols].view(t.rows, t.cols)
    assert torch.equal(view("dec.L1.wo"), sd["model.layers.1.self_attn.o_proj.weight"])
    wqkv = view("dec.L0.wqkv")
    assert torch.equal(wqkv[:256], sd["model.layers.0.self_attn.q_proj.weight"])
    assert torch.equal(wqkv[512:], sd["model.layers.0.self_attn.v_proj.weight"])
    wgu = view("dec.L0.wgu")  # interleaved (gate_i, up_i)
    assert torch.equal(wgu[0::2], sd["model.layers.0.mlp.gate_proj.weight"])
    assert torch.equal(wgu[1::2], sd["model.layers.0.mlp.up_proj.weight"])
    pw = view("vit.patch_w")
    assert torch.equal(pw[:, :588], sd["model.vision_model.vision_model.embeddings.patch_embedding.weight"].reshape(144, -1))
    assert pw[:, 588:].abs().sum() == 0
    v = "model.vision_model.vision_model.head."
    assert torch.equal(view("vit.head.wkv"), sd[v + "attention.in_proj_weight"][144:])


def test_invalid_config_is_rejected_not_crashing():
    from detikzify_b200 import _lib
    from detikzify_b200.engine import to_c_config
    from detikzify_b200.model.configuration import preset
    lib = _lib.load_library()
    cc = to_c_config(preset("tiny"))
    cc.head_dim = 64
    assert lib.dtk_weight_count(C.byref(cc)) < 0
    assert lib.dtk_arena_bytes(C.byref(cc)) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_product_path_fails_loudly_without_cuda():
    from detikzify_b200.engine import Engine, EngineError, pack_arena
    from detikzify_b200.model.configuration import preset
    from detikzify_b200.model.weights import random_init
    cfg = preset("tiny")
    with pytest.raises(EngineError, match="no CPU fallback"):
        Engine(cfg, pack_arena(cfg, random_init(cfg)), device=0)
    from detikzify_b200.model import load
    with pytest.raises(EngineError):
        load("tiny", device_map=0)


def test_plain_c_client_links_and_runs(tmp_path):
    """The boundary is a real C ABI: a C99 program (no CUDA / torch headers) compiles against include/detikzify_b200.h, links
    to the shared library and uses the host-only entry points (weight table, arena size, algorithmic decode bytes)."""
    import shutil
    import subprocess
    from pathlib import Path
    from detikzify_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None or not _lib.LIB_PATH.exists():
        pytest.skip("gcc or the built library is not available")
    root = Path(__file__).resolve().parents[1]
    exe = tmp_path / "abi_client"
    cmd = [gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", str(root / "include"), str(root / "tests" / "c" / "abi_client.c"),
           "-o", str(exe), "-L", str(_lib.LIB_PATH.parent), "-ldtk_b200", f"-Wl,-rpath,{_lib.LIB_PATH.parent}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok weights="), (r.returncode, r.stdout, r.stderr)
