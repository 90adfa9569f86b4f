#!/bin/bash
# End-of-session verification on a B200 box (run through gpurun): GPU test suite, both bench arms, the ncu launch list,
# ncu --set full captures of the kernels added this session, the stage-1 / encoder timings and the other BASELINE configs.
# Everything lands under gpurun_out/r1_s3_*; the summaries judged live under profiles/.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
(timeout 400 python -m pytest tests -q -m gpu 2>&1 | tail -4) > $O/r1_s3_tests.log
tail -2 $O/r1_s3_tests.log
timeout 300 python bench.py --steps 3 --warmup 3 > $O/r1_s3_bench.json 2> $O/r1_s3_bench.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $O/r1_s3_bench_reference_arm.json 2>> $O/r1_s3_bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r1_s3_launches.csv python tools/profile_step.py 2 > /dev/null 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:skinny -s 2 -c 1 -o $O/r1_s3_skinny_ff1 -f python tools/profile_skinny.py > /dev/null 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:skinny -s 5 -c 1 -o $O/r1_s3_skinny_ff2 -f python tools/profile_skinny.py > /dev/null 2>&1
timeout 100 python tools/profile_elementwise.py $O/r1_s3_elementwise.json > $O/r1_s3_elementwise.log 2>&1
timeout 150 ncu --set full --clock-control none -k regex:"add_noise_vec|cfg_ddim_step" -c 6 -o $O/r1_s3_elementwise2 -f python tools/profile_elementwise.py /dev/null > /dev/null 2>&1
timeout 150 python tools/bench_stage1.py $O/r1_s3_bench_stage1.json > $O/r1_s3_bench_stage1.log 2>&1
PCDM_ONLY_PRIOR=1 PCDM_PRIOR_FUSE_LN=1 timeout 100 python tools/bench_stage1.py $O/tmp_s1_ln.json 2>&1 | grep prior | cut -c1-140
timeout 250 python tools/bench_configs.py $O/r1_s3_bench_configs.json > $O/r1_s3_bench_configs.log 2>&1
cut -c1-330 $O/r1_s3_bench.json
tail -2 $O/r1_s3_bench.err
cut -c1-200 $O/r1_s3_bench_stage1.log
cat $O/r1_s3_elementwise.log
