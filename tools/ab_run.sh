#!/bin/bash
# Same-box A/B of library variants built with tools/build_variant.py into .ab/lib_<name>.so:
#   tools/ab_run.sh <name> ...   ("release" = the in-tree library); per variant: tools/dev_epi.py, tools/dev_gn_stats.py,
#   tools/step_time.py -> gpurun_out/ab_<tool>_<name>.log, then a side-by-side table.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out; mkdir -p $O
for v in "$@"; do
  if [ $v = release ]; then unset PCDM_B200_LIB; else export PCDM_B200_LIB=$PWD/.ab/lib_$v.so; fi
  timeout 200 python tools/dev_epi.py > $O/ab_epi_$v.log 2>&1
  [ -n "$AB_GN" ] && timeout 200 python tools/dev_gn_stats.py > $O/ab_gn_$v.log 2>&1
  timeout 200 python tools/step_time.py > $O/ab_step_$v.log 2>&1
  echo "== $v: $(tail -1 $O/ab_step_$v.log)"
done
for v in "$@"; do
  if [ $v = release ]; then unset PCDM_B200_LIB; else export PCDM_B200_LIB=$PWD/.ab/lib_$v.so; fi
  timeout 200 python tools/step_time.py > $O/ab_step2_$v.log 2>&1
  echo "== $v (second pass): $(tail -1 $O/ab_step2_$v.log)"
done
for v in "$@"; do echo "--- $v"; tail -n +2 $O/ab_epi_$v.log; [ -n "$AB_GN" ] && cat $O/ab_gn_$v.log; done
