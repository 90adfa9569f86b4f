#!/bin/bash
# End-of-round verification on a B200 box (run through gpurun): GPU test suite + smoke, both bench arms, the ncu launch
# list of one UNet step, ncu --set full captures of the round's kernels, the DRAM traffic of the dominant kernel class
# (-> profiles/roofline_traffic.json), the in-graph ablation, and the other BASELINE configurations.
# Everything lands under gpurun_out/r2_final_*; the summaries judged live under profiles/.   usage: final_run.sh [git head]
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
P=r2_final
HEAD=${1:-unknown}
rm -f $O/r2_parity.jsonl
(timeout 500 python -m pytest tests -q -m gpu 2>&1 | tail -4) > $O/${P}_tests.log
tail -2 $O/${P}_tests.log
python tools/collect_parity.py $O/r2_parity.jsonl $O/${P}_parity.json > $O/${P}_parity.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/${P}_smoke.log 2>&1; tail -3 $O/${P}_smoke.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/${P}_clocks.csv &
SMI=$!
timeout 300 python bench.py --steps 5 --warmup 3 > $O/${P}_bench.json 2> $O/${P}_bench.err
kill $SMI
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${P}_bench_reference_arm.json 2>> $O/${P}_bench.err
timeout 200 python tools/ablate.py $O/${P}_ablate.json > $O/${P}_ablate.log 2>&1
timeout 250 python tools/bench_configs.py $O/${P}_bench_configs.json > $O/${P}_bench_configs.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${P}_launches.csv python tools/profile_step.py 2 > /dev/null 2>&1
python tools/summarize_launches.py $O/${P}_launches.csv "Round 2 — launch list of one UNet evaluation" > $O/${P}_launches.md
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/${P}_convs.csv python tools/profile_convs.py > $O/${P}_convs.log 2>&1
python tools/extract_traffic.py $O/${P}_convs.csv $O/${P}_roofline_traffic.json $HEAD
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -o $O/${P}_kernels -f python tools/profile_r2.py > $O/${P}_ncu.log 2>&1
ncu -i $O/${P}_kernels.ncu-rep --page raw --csv > $O/${P}_kernels_raw.csv 2>/dev/null
cut -c1-400 $O/${P}_bench.json
tail -2 $O/${P}_bench.err
