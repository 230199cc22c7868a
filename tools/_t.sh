timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/prof_ops.py --ops dark4.1.dw,dark5.1.dw,up1.m0.g1.dw,up1.m0.g2.dw,dn2.m0.g1.dw --iters 4 2>&1 | tail -5
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; python - <<'PY'
import json
for l in open('gpurun_out/r2y_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_by_kind'])
PY
