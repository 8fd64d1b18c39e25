"""Recipe for ``oracle/_ref``: a verbatim, git-ignored copy of the reference's Python scripts (TEST INFRASTRUCTURE).

The reference is pure Python, so "building" it is copying its scripts next to the oracle; ``__graft_entry__.build()`` runs this in the
build container (where /root/reference exists).  ``oracle/_ref/`` is listed in .gitignore (the sources never enter this repository's
history) but not in .gpurunignore, so the copy travels to the GPU box, where ``bench.py --impl reference`` imports the UNMODIFIED
scripts through ``oracle/ref_harness.py`` and times the reference's own training / inference loops on the host cores.

    python oracle/build_ref.py
"""
import os
import shutil
import sys

SRC = "/root/reference"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
SCRIPTS = ("causal_anomaly_detection.py", "causal_anomaly_detection1.py", "minicausal_vad_complete3.py", "avenue_training_script2.py",
           "avenue_training_script_bbox.py", "json_utils.py")


def build(verbose: bool = True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"[oracle/build_ref] {SRC} is not visible here: keeping whatever {DST} already holds")
        return os.path.isdir(DST)
    os.makedirs(DST, exist_ok=True)
    for f in SCRIPTS:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    if verbose:
        print(f"[oracle/build_ref] copied {len(SCRIPTS)} reference scripts to {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
