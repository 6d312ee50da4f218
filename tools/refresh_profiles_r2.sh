#!/usr/bin/env bash
# Copies the round-2 measurement artefacts of `gpurun -- 'bash tools/gpu_refresh_r2.sh'` from gpurun_out/ (scratch) into
# profiles/ (tracked) and regenerates the text summaries of the ncu captures.  Run from the repo root.
set -euo pipefail
cd "$(dirname "${BASH_SOURCE[0]}")/.."
for f in r2_bench.json r2_bench_c2.json r2_bench_c4.json r2_bench_c5.json r2_bench_ref.json r2_smoke.log r2_pytest_gpu.log r2_env_throughput.txt \
         r2_e2e_sweep.txt r2_validation.txt r2_same_actions.log r2_block_tail.txt r2_binning_study.txt r2_contact_profile.txt r2_launches.csv \
         r2_scale_2.json r2_scale_4.json r2_scale_8.json; do
  [ -f gpurun_out/$f ] && cp gpurun_out/$f profiles/$f
done
[ -f profiles/r2_bench.json ] && cp profiles/r2_bench.json profiles/r2_scale_1.json
python tools/ncu_summary.py gpurun_out/r2_step_kernel.ncu-rep > profiles/r2_step_kernel_summary.txt 2>&1
python tools/ncu_regions.py gpurun_out/r2_step_kernel.ncu-rep > profiles/r2_step_kernel_regions.txt 2>&1
python tools/launch_summary.py gpurun_out/r2_launches.csv "ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 100 --csv  python bench.py --profile --steps 20 --warmup 3   (launches 500..600: the 64 action tensors and the reset take ~200 launches, a step takes 3, a timed step 4 -> warm-up and timed steps)" > profiles/r2_launches_summary.txt
python tools/ncu_summary.py gpurun_out/r2_env_kernels.ncu-rep 2>&1 | grep -E "^kernel|gpu__time_duration|dram__bytes|gpu__dram_throughput|sm__throughput|registers_per_thread|warps_active|block_size|grid_size|issue_active|stall" > profiles/r2_env_kernels_summary.txt
python - <<'PY'
import json, re
t = open('profiles/r2_step_kernel_summary.txt').read()
g = lambda k: float(re.search(re.escape(k) + r'\s+([0-9.]+)', t).group(1))
d = {"source": "ncu --set full --clock-control none, launch #110 (qg_step_kernel filter) of `python bench.py --profile --steps 20 --warmup 3` (100 pre-roll steps), B200, round 2",
     "envs": 65536, "frame_skip": 4, "gpu__time_duration_ms": g("gpu__time_duration.sum"),
     "dram__bytes_read.sum": g("dram__bytes_read.sum") * 1e6, "dram__bytes_write.sum": g("dram__bytes_write.sum") * 1e6,
     "smsp__inst_executed.sum": g("smsp__inst_executed.sum"),
     "active_lanes_per_instruction": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
     "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
     "registers_per_thread": g("launch__registers_per_thread")}
json.dump(d, open('profiles/r2_step_kernel_metrics.json', 'w'), indent=1)
print(d)
PY
