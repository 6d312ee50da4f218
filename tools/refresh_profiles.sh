#!/usr/bin/env bash
# Copies the measurement artefacts of the last gpurun calls from gpurun_out/ (scratch) into profiles/ (tracked) and
# regenerates the text summaries from the ncu captures.  Run from the repo root after the GPU calls listed in
# tools/README.md have written gpurun_out/{prof_final.ncu-rep, r1_launches.csv, r1_*.json, r1_*.log, r1_*.txt}.
set -euo pipefail
cd "$(dirname "${BASH_SOURCE[0]}")/.."
for f in r1_bench.json r1_bench_c2.json r1_bench_c4.json r1_bench_c5.json r1_bench_c5_8gpu.json r1_bench_ref.json r1_scale_2.json r1_scale_4.json \
         r1_scale_8.json r1_gpu_check.txt r1_pytest_gpu.log r1_smoke.log r1_validation.txt r1_same_actions.log r1_env_throughput.txt r1_launches.csv; do
  [ -f gpurun_out/$f ] && cp gpurun_out/$f profiles/$f
done
cp profiles/r1_bench.json profiles/r1_scale_1.json
python tools/ncu_summary.py gpurun_out/prof_final.ncu-rep > profiles/r1_step_kernel_summary.txt 2>&1
python tools/ncu_regions.py gpurun_out/prof_final.ncu-rep > profiles/r1_step_kernel_regions.txt 2>&1
python tools/launch_summary.py gpurun_out/r1_launches.csv "ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 100 --csv  python bench.py --profile --steps 20 --warmup 3   (launches 500..600: the 64 action tensors and the reset take ~200 launches, a step takes 3, a timed step 4 -> warm-up and timed steps)" > profiles/r1_launches_summary.txt
python - <<'PY'
import json, re
t = open('profiles/r1_step_kernel_summary.txt').read()
g = lambda k: float(re.search(re.escape(k) + r'\s+([0-9.]+)', t).group(1))
d = {"source": "ncu --set full --clock-control none, launch #110 (qg_step_kernel filter) of `python bench.py --profile --steps 20 --warmup 3` (100 pre-roll steps), B200",
     "envs": 65536, "frame_skip": 4, "gpu__time_duration_ms": g("gpu__time_duration.sum"),
     "dram__bytes_read.sum": g("dram__bytes_read.sum") * 1e6, "dram__bytes_write.sum": g("dram__bytes_write.sum") * 1e6,
     "smsp__inst_executed.sum": g("smsp__inst_executed.sum"),
     "active_lanes_per_instruction": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
     "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
     "registers_per_thread": g("launch__registers_per_thread")}
json.dump(d, open('profiles/r1_step_kernel_metrics.json', 'w'), indent=1)
print(d)
PY
