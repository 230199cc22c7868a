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
    ap.add_argument("--device", default="cpu", help="cuda: the same modules run eagerly by torch / cuDNN on the GPU (the 'library' bar)")
    ap.add_argument("--autocast", default="", help="bf16: torch.autocast(bfloat16) around the forward (cuda only)")
    ap.add_argument("--loss", type=int, default=0, help="N > 0: time the reference criterion (nets/yolo_training.py Loss) on "
                    "synthetic head maps with N targets per image instead of the forward path")
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
    dev = torch.device(a.device)
    on_gpu = dev.type == "cuda"
    net = net.to(dev)
    rgb, nir = torch.rand(a.batch, 3, a.size, a.size).to(dev), torch.rand(a.batch, 3, a.size, a.size).to(dev)
    dec = ns.DecodeBox(1, (a.size, a.size))
    shape = np.array([a.size, a.size])
    cast = torch.autocast("cuda", dtype=torch.bfloat16) if (on_gpu and a.autocast == "bf16") else None

    def forward():
        if cast is None:
            return net(rgb, nir)
        with cast:
            out = net(rgb, nir)
        return (out[0].float(), out[1].float(), [t.float() for t in out[2]], out[3].float(), out[4].float())

    def step():
        with torch.no_grad():
            out = forward()
            y = dec.decode_box(out)
            cand = int((y[..., 4:].max(-1)[0] >= 0.5).sum())
            res = dec.non_max_suppression(y, 1, [a.size, a.size], shape, True, conf_thres=0.5, nms_thres=0.3)
        return cand, res

    if a.loss:
        # the reference criterion on head maps of this size (validation step, utils/utils_fit_mul.py:78-92)
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from oracle import loss as OL
        from nets.yolo_training import Loss
        feats, targets = OL.synth_case(seed=105, B=a.batch, nc=1, hw0=(a.size // 8, a.size // 8), n_targets=a.loss)
        maps = [torch.from_numpy(f).to(dev) for f in feats]
        tgt = torch.from_numpy(targets).to(dev)
        crit = Loss(net)

        def step():   # noqa: F811
            with torch.no_grad():
                v = crit((maps[0], maps[0], maps), tgt)
            return 0, [float(v)]

    def sync():
        if on_gpu:
            torch.cuda.synchronize()

    for _ in range(a.warmup):
        step()
    times, cand = [], 0
    sync()
    t_start = time.perf_counter()
    for _ in range(a.steps):
        t0 = time.perf_counter()
        cand, res = step()
        sync()
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
        "device": a.device, "autocast": a.autocast, "loss_targets_per_image": a.loss,
        "gpu": torch.cuda.get_device_name(0) if on_gpu else None,
        "nms_candidates": cand // a.batch, "kept": [0 if r is None else (r if isinstance(r, float) else int(len(r))) for r in res],
        "torch": torch.__version__}))


if __name__ == "__main__":
    main()
