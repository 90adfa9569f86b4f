cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out; mkdir -p $O
(timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -6) > $O/s3_tests.log; tail -3 $O/s3_tests.log
for v in release base gelu prefetch; do
  if [ $v = release ]; then unset PCDM_B200_LIB; else export PCDM_B200_LIB=$PWD/.ab/lib_$v.so; fi
  timeout 200 python tools/dev_epi.py > $O/s3_epi_$v.log 2>&1
  timeout 200 python tools/step_time.py > $O/s3_step_$v.log 2>&1
  echo "== $v"; tail -1 $O/s3_step_$v.log
done
unset PCDM_B200_LIB
timeout 200 python tools/step_time.py > $O/s3_step_release2.log 2>&1; tail -1 $O/s3_step_release2.log
paste $O/s3_epi_base.log $O/s3_epi_gelu.log $O/s3_epi_prefetch.log $O/s3_epi_release.log | awk -F'\t' '{print $1 " | " substr($2,43) " | " substr($3,43) " | " substr($4,43)}'
