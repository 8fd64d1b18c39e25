#!/bin/bash
# Build libcvad_b200.so from another git ref into variants/<name>.so (git-ignored, travels to the GPU box) so one gpurun call can
# A/B it against the in-tree library:  CVAD_B200_LIB=variants/<name>.so python -m pytest ... / python bench.py ...
#   tools/build_variant.sh <git-ref> <name>
set -e
ref=$1; name=$2
root=$(cd "$(dirname "$0")/.." && pwd)
wt=$(mktemp -d /tmp/cvad_variant.XXXXXX)
git -C "$root" worktree add --detach "$wt" "$ref" >/dev/null 2>&1
pkg="$wt/causal-learning-based-video-anomaly-detection_paper_code_raw_b200"
mkdir -p "$root/variants" "$wt/obj"
for f in "$pkg"/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
       -I "$wt/include" -I "$pkg/csrc" -c "$f" -o "$wt/obj/$(basename "${f%.cu}").o" &
done
wait
nvcc -shared --cudart shared -o "$root/variants/$name.so" "$wt"/obj/*.o
git -C "$root" worktree remove --force "$wt"
echo "built variants/$name.so from $ref"
