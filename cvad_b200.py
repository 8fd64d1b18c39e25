"""Import alias: registers ``causal-learning-based-video-anomaly-detection_paper_code_raw_b200/`` as package ``cvad_b200``.

    import cvad_b200
    from cvad_b200.mb import ImprovedMiniCausalVAD
"""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "causal-learning-based-video-anomaly-detection_paper_code_raw_b200")
_spec = importlib.util.spec_from_file_location("cvad_b200", os.path.join(_DIR, "__init__.py"),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["cvad_b200"] = _mod
_spec.loader.exec_module(_mod)
