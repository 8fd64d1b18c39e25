"""SASS opcode evidence for libcvad_b200.so: per kernel, how many tcgen05 MMAs (UTCHMMA = f16/bf16/tf32 kinds), TMA loads/stores
(UTMALDG / UTMASTG), TMEM loads/stores (LDTM / STTM), tensor-core barriers (UTCBAR) and mbarrier ops (SYNCS) the compiled code holds.

    python tools/sass_summary.py > profiles/r02_sass_opcodes.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "causal-learning-based-video-anomaly-detection_paper_code_raw_b200", "libcvad_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCATOMSWS", "SYNCS", "HMMA", "FFMA", "RED", "ATOMG"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur.replace("(anonymous namespace)::", "")).replace("void ", "").strip()
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for o in OPS:
                if op == o or op.startswith(o + "."):
                    per[cur][o] += 1
    arch = [a for a in re.findall(r"arch = (sm_\w+)", out) if a.startswith("sm_1")]
    print(f"# SASS opcode counts per kernel of libcvad_b200.so ({', '.join(sorted(set(arch)))}; `cuobjdump -sass`, tools/sass_summary.py)\n")
    print("UTCHMMA = tcgen05.mma (kind::f16 / kind::tf32), UTMALDG = TMA tile load, LDTM = tcgen05.ld (TMEM -> registers), UTCBAR = tcgen05.commit,")
    print("SYNCS = mbarrier arrive/wait.  Kernels without tensor-core or TMA opcodes (bandwidth / latency kernels) are summarised in the last row.\n")
    print("| kernel | " + " | ".join(OPS) + " |")
    print("|---|" + "---:|" * len(OPS))
    tot, rest, nrest = collections.Counter(), collections.Counter(), 0
    for k, c in per.items():
        tot.update(c)
        if not any(c[o] for o in ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR")):
            rest.update(c)
            nrest += 1
            continue
        print(f"| `{k}` | " + " | ".join(str(c[o]) if c[o] else "" for o in OPS) + " |")
    print(f"| *{nrest} other kernels (no tensor-core / TMA opcodes)* | " + " | ".join(str(rest[o]) if rest[o] else "" for o in OPS) + " |")
    print(f"| **total ({len(per)} kernels)** | " + " | ".join(str(tot[o]) if tot[o] else "" for o in OPS) + " |")


if __name__ == "__main__":
    sys.exit(main())
