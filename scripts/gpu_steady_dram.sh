#!/bin/bash
# Steady-state DRAM bytes per FK launch at the contract batch: ONE-pass metrics (no kernel replay, caches untouched) on the
# launches of bench.py's rotation over 235 MB of buffers.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none \
    -k regex:fk_jacobian_kernel -s 80 -c 64 --csv --log-file gpurun_out/steady_dram.csv \
    python bench.py --steps 200 --warmup 3 --reps 1 --no-cpu-baseline --no-e2e --no-large --no-modes --no-sharded > gpurun_out/steady_dram.log 2>&1
python - <<'PY'
import csv, json
rows = [r for r in csv.DictReader(l for l in open("gpurun_out/steady_dram.csv") if not l.startswith("=="))]
acc = {}
for r in rows:
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "us": 1, "ns": 1e-3, "ms": 1e3}.get(u, 1)
    acc.setdefault(r["Metric Name"], []).append(v)
n = len(acc["dram__bytes_read.sum"])
out = {"launches": n, "dram_bytes_read_per_launch": sum(acc["dram__bytes_read.sum"]) / n,
       "dram_bytes_write_per_launch": sum(acc["dram__bytes_write.sum"]) / n,
       "gpu_time_us_per_launch_under_ncu": sum(acc["gpu__time_duration.sum"]) / n, "algorithmic_bytes_per_launch": 65536 * 224,
       "how": "ncu one-pass metrics (no replay), --cache-control none, launches 80..143 of bench.py's rotation over 16 buffer sets (235 MB > L2)"}
out["dram_bytes_per_launch_batch65536"] = out["dram_bytes_read_per_launch"] + out["dram_bytes_write_per_launch"]
json.dump(out, open("gpurun_out/steady_dram.json", "w"), indent=1)
print(json.dumps(out))
PY
