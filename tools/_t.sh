timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "conv or dfl" 2>&1 | tail -3
timeout 300 python tools/prof_ops.py --ops dn1.cv1,dn1.cv2,up1.cv2,sppf.cv1,up2.cv2,up2.cv1,head0.out,head1.out,dark4.0,head0.0,dark2.0,up1.m0.g1.pw --iters 4 2>&1 | tail -12
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; python - <<'PY'
import json
for l in open('gpurun_out/r2z_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['ms_by_kind'])
PY
