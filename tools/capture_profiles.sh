#!/bin/bash
# The round-2 evidence capture (run on the GPU box: tools/gpu_retry.sh 2400 -- "bash tools/capture_profiles.sh"): GPU test suite, one
# default bench line, the ncu launch list of the bench command, and one --set full capture of a whole step (tools/ncu_step.py)
# exported as CSV; tools/ncu_join.py turns the raw CSV + unit manifest into profiles/r2_step_full.csv and conv_traffic.json.
set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/cap_pytest.log 2>&1; tail -3 gpurun_out/cap_pytest.log
timeout 600 python bench.py > gpurun_out/cap_bench_n1.json 2> gpurun_out/cap_bench_n1.err; tail -c 600 gpurun_out/cap_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-count 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/cap_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o /tmp/r2_step python tools/ncu_step.py --manifest gpurun_out/r2_step_units.json > gpurun_out/cap_ncu_step.log 2>&1; tail -2 gpurun_out/cap_ncu_step.log
ls -la /tmp/r2_step.ncu-rep
ncu -i /tmp/r2_step.ncu-rep --page raw --csv > gpurun_out/r2_step_raw.csv 2>/dev/null
ncu -i /tmp/r2_step.ncu-rep --page source --csv --kernel-name regex:stem_kernel > gpurun_out/r2_stem_source.csv 2>/dev/null
ncu -i /tmp/r2_step.ncu-rep --page source --csv --kernel-name regex:chain_kernel --launch-skip 0 --launch-count 1 > gpurun_out/r2_chain_source.csv 2>/dev/null
ncu -i /tmp/r2_step.ncu-rep --page source --csv --kernel-name regex:cbam --launch-skip 0 --launch-count 1 > gpurun_out/r2_cbam_source.csv 2>/dev/null
sz=$(stat -c %s /tmp/r2_step.ncu-rep); if [ "$sz" -lt 40000000 ]; then cp /tmp/r2_step.ncu-rep gpurun_out/; fi
ls -la gpurun_out | tail -12
