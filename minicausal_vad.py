"""Drop-in for the ``minicausal_vad`` module that avenue_training_script1.py imports (s1:20) and the reference does not ship.

    from minicausal_vad import MiniCausalVAD

The class is the B200-native trainer of the checkpointed model (M-B, avenue_training_script2.py:69-297 minus the
``_improved`` suffixes); all arithmetic runs in libcvad_b200.so."""
import cvad_b200  # noqa: F401  (registers the package; raises if the CUDA extension is missing)
from cvad_b200.mb import CausalAnomalyDetector, ImprovedMiniCausalVAD, MiniCausalVAD  # noqa: F401

__all__ = ["MiniCausalVAD", "ImprovedMiniCausalVAD", "CausalAnomalyDetector"]
