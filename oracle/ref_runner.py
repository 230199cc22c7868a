"""TEST INFRASTRUCTURE / CPU BASELINE -- times the REAL reference's hot path on the host cores, in its own process
(the reference's `nets` / `utils` packages clash with the drop-in's).  Called by bench.py:

    python oracle/ref_runner.py --phi n --size 640 --batch 1 --steps 30 --warmup 5 [--seconds S]

Methodology = BASELINE.md 4 / the reference's own get_FPS loop (yolo_mul.py:132-166): net(rgb, depth) ->
decode_box -> non_max_suppression(conf 0.5, IoU 0.3) per iteration, model.eval(), torch.no_grad(), constructor
init (weights_init N(0, 0.02)), torch.manual_seed(0) inputs, all host threads.  phi='n' at 640x640 runs the
reference exactly as shipped; any other configuration needs the five-constant generalisation (oracle/ref_model.py).
Prints one JSON object.
"""
import argparse
import json
import os
import platform
import sys
import time
import warnings

warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor() or "unknown"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--phi", default="n")
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--seconds", type=float, default=0.0, help="stop the timed loop after this many seconds (0: run all steps)")
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args()
    import numpy as np
    import torch
    import ref_model
    ns = ref_model.load()
    threads = a.threads or (os.cpu_count() or 1)
    torch.set_num_threads(threads)
    shipped = a.phi == 'n' and a.size == 640
    torch.manual_seed(0)
    net = ref_model.build(ns, a.phi, a.size, a.size, 1, shipped)
    rgb, nir = torch.rand(a.batch, 3, a.size, a.size), torch.rand(a.batch, 3, a.size, a.size)
    dec = ns.DecodeBox(1, (a.size, a.size))
    shape = np.array([a.size, a.size])

    def step():
        with torch.no_grad():
            out = net(rgb, nir)
            y = dec.decode_box(out)
            cand = int((y[..., 4:].max(-1)[0] >= 0.5).sum())
            res = dec.non_max_suppression(y, 1, [a.size, a.size], shape, True, conf_thres=0.5, nms_thres=0.3)
        return cand, res

    for _ in range(a.warmup):
        step()
    times, cand = [], 0
    t_start = time.perf_counter()
    for _ in range(a.steps):
        t0 = time.perf_counter()
        cand, res = step()
        times.append(time.perf_counter() - t0)
        if a.seconds and time.perf_counter() - t_start >= a.seconds:
            break
    total = sum(times)
    print(json.dumps({
        "kind": "reference", "root": "oracle/_ref" if ns.root.endswith("_ref") else ns.root, "shipped_unmodified": shipped,
        "phi": a.phi, "size": a.size, "batch": a.batch, "steps": len(times), "warmup": a.warmup,
        "pairs_per_s": a.batch * len(times) / total, "pairs_per_s_best": a.batch / min(times),
        "ms_median": 1e3 * float(np.median(times)), "ms_best": 1e3 * min(times), "seconds": total,
        "threads": torch.get_num_threads(), "cpu_count": os.cpu_count(), "cpu_model": cpu_model(),
        "nms_candidates": cand // a.batch, "kept": [0 if r is None else int(len(r)) for r in res],
        "torch": torch.__version__}))


if __name__ == "__main__":
    main()
