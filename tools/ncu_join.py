#!/usr/bin/env python
"""Join an `ncu --page raw --csv` dump of tools/ncu_step.py with its unit manifest: one row per kernel launch with the
plan unit it belongs to, time, DRAM bytes against the unit's algorithmic bytes, and pipe utilisation.

    python tools/ncu_join.py gpurun_out/r2_step_raw.csv gpurun_out/r2_step_units.json profiles/r2_step_full.csv \
        [profiles/conv_traffic.json]
"""
import csv
import json
import sys

raw, manifest, out = sys.argv[1:4]
rows = list(csv.reader(open(raw)))
hdr, body = rows[0], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
foreign = [r for r in body if "at::" in r[col["Kernel Name"]]]       # torch's own kernels (e.g. a counter fill): not plan units
body = [r for r in body if "at::" not in r[col["Kernel Name"]]]
units = json.load(open(manifest))
launches = []
for u in units["units"]:
    for k in range(u["launches"]):
        launches.append((u, k))
assert len(launches) == len(body), (len(launches), len(body))


def num(r, name, default=0.0):
    try:
        return float(r[col[name]].replace(",", ""))
    except (KeyError, ValueError):
        return default


def unit_of(name):
    return rows[1][col[name]]


assert unit_of("gpu__time_duration.sum") == "us" and unit_of("dram__bytes_read.sum") in ("Mbyte", "Kbyte", "byte", "Gbyte")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
fields = ["launch", "unit", "kind", "kernel", "grid", "block", "regs", "time_us", "dram_read_MB", "dram_write_MB", "algorithmic_MB",
          "dram_over_algorithmic", "gpu__dram_throughput_pct", "sm__mem_tensor_cycles_active_pct", "sm__issue_active_pct",
          "lsu_wavefronts_shared_pct", "sm__warps_active_pct"]
w = csv.writer(open(out, "w"))
w.writerow(fields)
tot = {}
conv_rd = conv_wr = 0.0
conv_n = 0
for i, (r, (u, k)) in enumerate(zip(body, launches)):
    kern = r[col["Kernel Name"]]
    kern = kern.split("(")[0].replace("void ", "").replace("dcfa::<unnamed>::", "")
    rd = num(r, "dram__bytes_read.sum") * scale[unit_of("dram__bytes_read.sum")]
    wr = num(r, "dram__bytes_write.sum") * scale[unit_of("dram__bytes_write.sum")]
    t = num(r, "gpu__time_duration.sum")
    alg = u["bytes"] if k == 0 else 0
    w.writerow([i, u["name"], u["kind"], kern, r[col["Grid Size"]], r[col["Block Size"]], r[col.get("launch__registers_per_thread", 0)],
                "%.2f" % t, "%.2f" % (rd / 1e6), "%.2f" % (wr / 1e6), "%.2f" % (alg / 1e6),
                "%.2f" % ((rd + wr) / alg) if alg else "",
                "%.1f" % num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                "%.1f" % num(r, "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                "%.1f" % num(r, "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
                "%.1f" % num(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
                "%.1f" % num(r, "sm__warps_active.avg.pct_of_peak_sustained_active")])
    a = tot.setdefault(u["kind"], [0, 0.0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += t; a[2] += rd; a[3] += wr; a[4] += alg
    if u["kind"] == "conv":
        conv_rd += rd; conv_wr += wr; conv_n += 1
print("%-10s %4s %10s %10s %10s %10s" % ("kind", "n", "time_us", "read_MB", "write_MB", "algo_MB"))
for k, a in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-10s %4d %10.1f %10.1f %10.1f %10.1f" % (k, a[0], a[1], a[2] / 1e6, a[3] / 1e6, a[4] / 1e6))
for r in foreign:
    print("foreign kernel: %s  %.2f us" % (r[col["Kernel Name"]][:80], float(r[col["gpu__time_duration.sum"]])))
print("total time_us %.1f over %d launches" % (sum(a[1] for a in tot.values()), len(body)))
if len(sys.argv) > 4:
    json.dump({"kernel": "conv_tma_kernel + conv_strip_kernel", "phi": units["phi"], "batch": units["batch"], "size": units["size"],
               "launches_per_step": conv_n, "dram_bytes_read_per_step": int(conv_rd), "dram_bytes_write_per_step": int(conv_wr),
               "source": "%s (ncu --set full --clock-control none over one step launched by tools/ncu_step.py, one launch per "
                         "layer, cold caches) summed over the %d records bench.py times as 'conv'" % (out, conv_n)},
              open(sys.argv[4], "w"), indent=1)
