timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --no-cpu-baseline --profile-ops gpurun_out/r2u_ops.json > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; tail -c 400 gpurun_out/r2u_bench.json
