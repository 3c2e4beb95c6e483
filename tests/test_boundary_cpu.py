"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares, the
nn.Module mirror has the reference's constructor / state_dict contract, and the product path refuses to run without CUDA."""
import os
import re

import numpy as np
import pytest
import torch

import dualvgr_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import dualvgr_videoqa_b200._lib as L
    header = open(os.path.join(ROOT, "include", "dualvgr_b200.h")).read()
    declared = set(re.findall(r"\b(dvgr_[a-z0-9_]+)\s*\(", header))
    declared -= {"dvgr_operand", "dvgr_gemm_args", "dvgr_lstm_args", "dvgr_lstm_seq_args", "dvgr_seg", "dvgr_wgrad_problem", "dvgr_colsum_problem", "dvgr_gat_args", "dvgr_gat_graph", "dvgr_lstm32_args"}
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(L.lib, name), f"{name} declared in include/dualvgr_b200.h but not exported"
    assert set(L.EXPORTED) == declared
    assert L.lib.dvgr_abi_version() == 1


def test_struct_layouts_match_header_sizes():
    """ctypes mirrors must have the C layout (sizes computed from the header's field lists with natural alignment)."""
    import ctypes
    import dualvgr_videoqa_b200._lib as L
    assert ctypes.sizeof(L.Operand) == 8 + 4 + 4 + 32 + 32
    assert ctypes.sizeof(L.GemmArgs) % 8 == 0
    assert ctypes.sizeof(L.GatGraph) == 8 * 10 + 8            # 10 pointers + uint (padded)
    assert ctypes.sizeof(L.GatArgs) == 4 * ctypes.sizeof(L.GatGraph) + 5 * 4 + 4 + 8 * 3 + 4 * 3 + 4 + 8


@pytest.mark.parametrize("name", ["g1_B4_N8_U2", "g2_B3_N20_U3", "g3_B5_N16_U1", "g4_B16_N20_U3"])
def test_state_dict_contract_matches_reference(golden, name):
    import dualvgr_videoqa_b200.model.models as M
    g = golden(name)
    B, N, L, A, V, U = [int(x) for x in g["cfg"]]
    m = M.DualVGR(vocab=orc.make_vocab(V, A), num_of_nodes=N, graph_module="GAT", graph_layers=1, unit_layers=U)
    sd = m.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["sd_keys"]]                      # same keys, same ORDER
    assert [",".join(map(str, v.shape)) for v in sd.values()] == [str(s) for s in g["sd_shapes"]]
    m.load_state_dict(orc.make_state_dict(U, A, V), strict=True)                  # validate.py:286 uses strict loading
    # adjacency: same values as the reference's scipy construction, and NOT part of the state_dict
    assert torch.allclose(m.visual_input_unit.appearance_adj, orc.build_adjacency(N))
    assert not any("adj" in k for k in sd)


def test_constructor_defaults_and_init_match_reference_conventions():
    import dualvgr_videoqa_b200.model.models as M
    import inspect
    sig = inspect.signature(M.DualVGR.__init__)
    assert [(k, v.default) for k, v in list(sig.parameters.items())[1:]] == [
        ("vision_dim", 2048), ("module_dim", 768), ("word_dim", 300), ("vocab", None), ("num_of_nodes", 8),
        ("graph_module", "GCN"), ("graph_layers", 1), ("unit_layers", 2)]
    torch.manual_seed(0)
    m = M.DualVGR(vocab=orc.make_vocab(12, 5), num_of_nodes=8, graph_module="GAT", unit_layers=1)
    # init_modules: every Linear / LSTM bias is zero, embedding is U(-1, 1)  (reference model/models.py:52-53)
    for n, p in m.named_parameters():
        if "bias" in n and "classifier.3" not in n:
            assert float(p.abs().max()) == 0.0, n
    e = m.linguistic_input_unit.encoder_embed.weight
    assert float(e.min()) >= -1 and float(e.max()) <= 1 and float(e.abs().mean()) > 0.3


def test_product_path_has_no_cpu_fallback():
    import dualvgr_videoqa_b200.model.models as M
    import dualvgr_videoqa_b200.ops as ops
    from dualvgr_videoqa_b200._lib import DvgrError
    m = M.DualVGR(vocab=orc.make_vocab(12, 5), num_of_nodes=8, graph_module="GAT", unit_layers=1)
    app, mot, q, qlen, _ = orc.make_inputs(2, 8, 5, 5, 12)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(app, mot, q, qlen)
    with pytest.raises(DvgrError):
        ops.linear_fwd(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "dualvgr-videoqa_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "dualvgr_oracle" not in src and "ref_shim" not in src, os.path.join(dirpath, f)


def test_model_importable_as_top_level_model_package():
    """train.py does `import model.models as modelset` with the package directory on sys.path."""
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); import model.models as modelset; "
            "print(modelset.DualVGR.__module__)" % os.path.join(ROOT, "dualvgr-videoqa_b200"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == "model.models"
