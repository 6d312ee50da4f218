"""Summarise an .ncu-rep (raw page + source page hot spots) into text for profiles/."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum', 'l1tex__t_sector_hit_rate.pct', 'smsp__sass_thread_inst_executed_op_fp32_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum', 'sm__sass_thread_inst_executed_op_fp64_pred_on.sum']
for r in rows[2:]:
    d = dict(zip(h, r))
    print("kernel:", d.get("Kernel Name", "")[:80])
    for k in keys:
        if k in d:
            print(f"  {k:70s} {d[k]}  {rows[1][h.index(k)]}")
    stalls = {k: float(v) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and v}
    if not stalls:
        stalls = {k: float(v) for k, v in d.items() if "issue_stalled" in k and k.endswith(".pct") and v}
    for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:10]:
        print(f"  stall {k:80s} {v:.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = None
for i, r in enumerate(rows):
    if "Source" in r and any("Instructions Executed" in c for c in r):
        hi = i
        break
if hi is not None:
    h = rows[hi]
    ci = h.index("Source")
    ce = [i for i, c in enumerate(h) if c == "Instructions Executed"][0]
    cs = [i for i, c in enumerate(h) if c.startswith("Warp Stall Sampling (All")]
    tot = 0
    items = []
    for r in rows[hi + 1:]:
        if len(r) != len(h):
            continue
        try:
            e = float(r[ce] or 0)
            s = float(r[cs[0]] or 0) if cs else 0
        except ValueError:
            continue
        items.append((s, e, r[ci][:110], r[0]))
    tots = sum(x[0] for x in items) or 1
    tote = sum(x[1] for x in items) or 1
    print("\ntop source lines by stall samples (share of samples, share of warp-instructions executed):")
    for s, e, txt, ln in sorted(items, key=lambda x: -x[0])[:45]:
        print(f"  {100*s/tots:5.1f}%  {100*e/tote:5.1f}%  L{ln}: {txt}")
