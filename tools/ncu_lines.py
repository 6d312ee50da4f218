"""Attribute an ncu source-page capture to CUDA source lines by joining it (by instruction order) with
nvdisasm --print-line-info of the kernel in libquadgym.so.  Usage: ncu_lines.py rep.ncu-rep [kernel-substr]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep = sys.argv[1]
ksub = sys.argv[2] if len(sys.argv) > 2 else "qg_step_kernelILb0ELi0"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "quadruped_gym_b200", "libquadgym.so")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
cur, fn, lines = None, None, []
for line in dis.split("\n"):
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\.text\.(\S+):", line)
    if m:
        fn = m.group(1)
        continue
    if fn and ksub in fn and re.match(r"\s+/\*[0-9a-f]+\*/\s+\S", line):
        lines.append((cur, line.split("*/", 1)[1].strip()[:60]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if "Source" in r and any("Instructions Executed" in c for c in r)][0]
h = rows[hi]
ce = h.index("Instructions Executed")
cs = [i for i, c in enumerate(h) if c.startswith("Warp Stall Sampling (All")][0]
ct = [i for i, c in enumerate(h) if c.startswith("Thread Instructions Executed")]
body = [r for r in rows[hi + 1:] if len(r) == len(h)]
print("sass instructions: nvdisasm", len(lines), "ncu", len(body))
n = min(len(lines), len(body))
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for i in range(n):
    (f, l) = lines[i][0] or ("?", 0)
    e = float(body[i][ce] or 0); s = float(body[i][cs] or 0); t = float(body[i][ct[0]] or 0) if ct else 0
    a = agg[(f, l)]
    a[0] += e; a[1] += s; a[2] += t
te = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print(f"total warp-instr {te:.3e} samples {ts:.0f}")
# buckets of 10 source lines
b = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for (f, l), a in agg.items():
    k = (f, l // 10 * 10)
    for j in range(3):
        b[k][j] += a[j]
print("share_instr share_stall lanes  file:lines")
for (f, l), a in sorted(b.items(), key=lambda kv: -kv[1][1])[:int(os.environ.get("NCU_TOP","45"))]:
    print(f"  {100*a[0]/te:5.1f}%  {100*a[1]/ts:5.1f}%  {a[2]/max(a[0],1):5.1f}  {f}:{l}-{l+9}")
