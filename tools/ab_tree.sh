#!/bin/bash
# Same-box A/B of whole trees: materialise commit $1 under .ab/<short>/ (git-ignored, but shipped to the GPU box by
# gpurun) with its own library built, so that one gpurun call can time `python tools/step_time.py` in both trees:
#   tools/ab_tree.sh <commit>;  gpurun -- 'python tools/step_time.py; (cd .ab/<short> && python tools/step_time.py)'
set -e
cd "$(dirname "$0")/.."
short=$(git rev-parse --short "$1")
rm -rf ".ab/$short"
mkdir -p ".ab/$short"
git archive "$1" | tar -x -C ".ab/$short"
(cd ".ab/$short" && python -m pcdms_b200.build > /dev/null && rm -rf pcdms_b200/build)
echo ".ab/$short"
