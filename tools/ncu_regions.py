"""Attribute an ncu capture to top-level source lines of physics_step / the step kernel using the inline
chains of nvdisasm --print-line-info-inline.  Usage: ncu_regions.py rep.ncu-rep"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep = sys.argv[1]
ksub = os.environ.get("NCU_KSUB", "qg_step_kernelILb0ELi0")
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.environ.get("QG_LIB", os.path.join(root, "quadruped_gym_b200", "libquadgym.so"))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
fn, chain, last, lines = None, [], None, []
for line in dis.split("\n"):
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
    if m:
        chain.append((m.group(1).split("/")[-1], int(m.group(2)), m.group(3).split("/")[-1] if m.group(3) else None))
        continue
    m = re.match(r"\.text\.(\S+):", line)
    if m:
        fn = m.group(1); continue
    if fn and ksub in fn and re.match(r"\s+/\*[0-9a-f]+\*/\s+\S", line):
        if chain:
            last = chain
        chain = []
        key = ("?", 0)
        if last:
            # frame directly inside the kernel body: prefer the physics_step-level line (inlined at qg_kernels)
            key = (last[-1][0], last[-1][1])
            for f, l, at in last:
                if f == "qg_step.cuh" and at == "qg_kernels.cuh":
                    key = (f, l)
        lines.append(key)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if "Source" in r and any("Instructions Executed" in c for c in r)][0]
h = rows[hi]
ce = h.index("Instructions Executed"); cs = [i for i, c in enumerate(h) if c.startswith("Warp Stall Sampling (All")][0]
ct = [i for i, c in enumerate(h) if c.startswith("Thread Instructions Executed")][0]
REASONS = ["stall_barrier", "stall_wait", "stall_short_sb", "stall_long_sb", "stall_branch_resolving", "stall_no_inst",
           "stall_not_selected", "stall_selected", "stall_dispatch", "stall_math", "stall_mio", "stall_lg"]
cr = [h.index(r) for r in REASONS]
body = [r for r in rows[hi + 1:] if len(r) == len(h)]
assert len(body) == len(lines), (len(body), len(lines))
def _mark(txt):
    src_lines = open(os.path.join(root, "quadruped_gym_b200", "csrc", "qg_step.cuh")).read().split("\n")
    for i, l in enumerate(src_lines):
        if txt in l:
            return i + 1
    raise SystemExit("marker not found: " + txt)

_M = [("base frame + FK", "DI void physics_step("), ("collision", "collide_lane(P, verts"), ("dynamics (RNE+CRB+act)", "// ---- velocity recursion and inertias"),
      ("rows setup", "// ---- constraint rows: joint limits"), ("phase: init", "// ---- phase machine around ONE arrow solve"),
      ("phase: arrow_solve", "arrow_solve(Hll, Hbl, Hc, Mbb"), ("phase: warmstart", "if (phase == 2) { done = true; continue; }"),
      ("phase: linesearch", "is the Newton direction: exact line search"), ("phase: update", "float gb[6], gl[3];"),
      ("phase: implicit prep", "// ---- next solve: implicit integration"), ("phase: hessian", "// ---- next solve: Newton direction"),
      ("sensors+integrate", "// ---- sensors of this forward pass")]
_L = [(n, _mark(t)) for n, t in _M]
REG = [(n, "qg_step.cuh", l, (_L[i + 1][1] if i + 1 < len(_L) else 100000)) for i, (n, l) in enumerate(_L)]
REG.append(("kernel prologue/epilogue", "qg_kernels.cuh", 0, 100000))
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0])
rs = collections.defaultdict(lambda: [0.0] * len(REASONS))
for key, r in zip(lines, body):
    name = "other " + key[0]
    for n, f, a, b in REG:
        if key[0] == f and a <= key[1] < b:
            name = n; break
    x = agg[name]
    x[0] += float(r[ce] or 0); x[1] += float(r[cs] or 0); x[2] += float(r[ct] or 0); x[3] += 1
    for j, c in enumerate(cr):
        rs[name][j] += float(r[c] or 0)
te = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print(f"total warp-instr {te:.3e}; stall samples {ts:.0f}")
print(f"{'region':28s} {'sass':>6s} {'instr%':>7s} {'time%':>7s} {'lanes':>6s}")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:28s} {a[3]:6d} {100*a[0]/te:6.1f}% {100*a[1]/ts:6.1f}% {a[2]/max(a[0],1):6.1f}")
print()
print("stall mix per region (% of the region's samples): " + " ".join(r.replace("stall_", "")[:7] for r in REASONS))
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    t = max(sum(rs[n]), 1.0)
    print(f"{n:28s} " + " ".join(f"{100 * v / t:7.0f}" for v in rs[n]))
