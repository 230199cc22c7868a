timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "sppf or cbam" 2>&1 | tail -15
timeout 300 python tools/prof_ops.py --ops sppf.cbam1.pool+19 --iters 4 2>&1 | tail -2
DCFA_SPPF_FUSED=0 timeout 300 python tools/prof_ops.py --ops sppf.cbam1.pool+19 --iters 4 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; tail -c 1500 gpurun_out/r2t_bench.json
