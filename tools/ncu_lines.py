"""Aggregate an ncu SASS source page by CUDA source line (ncu's CLI prints metrics only on the SASS view).
   usage: ncu_lines.py <report.ncu-rep> <launch index> <cubin> [top]
Joins per-instruction 'Instructions Executed' / stall samples with nvdisasm --print-line-info by offset."""
import csv, io, re, subprocess, sys

rep, idx, cubin = sys.argv[1], int(sys.argv[2]), sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kname = rows[0][1]
h = rows[1]
ai, ci, wi, si = h.index("Address"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)"), h.index("Source")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
inst = []
for r in rows[2:]:
    try:
        inst.append((int(r[ai], 16), int(r[ci]), int(r[wi] or 0), r[si].strip(), [int(r[i] or 0) for i in stall_cols]))
    except Exception:
        pass
base = inst[0][0]
mangled = None
# find the mangled name through cuobjdump symbol listing: match template args loosely
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
# split per function
funcs = re.split(r"\n\s*\.section\s+\.text\.", dis)
short = re.sub(r"\(.*", "", kname).replace("void ", "").split("::")[-1].split("<")[0]
targs = re.findall(r"\)(-?\d+)|\(bool\)(\d)", kname)
cands = [f for f in funcs if short in f.split("\n", 1)[0]]
def score(f):
    name = f.split("\n", 1)[0]
    digits = re.findall(r"ILi(\d+)|Lb(\d)", name)
    return name
sel = None
want = "".join(a or b for a, b in targs)
for f in cands:
    name = f.split("\n", 1)[0]
    got = "".join(a or b for a, b in re.findall(r"Li(\d+)E|Lb(\d)E", name))
    if got == want:
        sel = f
        break
if sel is None:
    sel = cands[0]
line = None
off2line = {}
for l in sel.split("\n"):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m and line:
        off2line[int(m.group(1), 16)] = line
agg = {}
tot = sum(i[1] for i in inst)
tw = sum(i[2] for i in inst)
for a, c, w, s, st in inst:
    ln = off2line.get(a - base, ("?", 0))
    e = agg.setdefault(ln, [0, 0, [0] * len(stall_cols)])
    e[0] += c
    e[1] += w
    e[2] = [x + y for x, y in zip(e[2], st)]
src = {}
print(kname[:100])
print("total warp-instructions", tot, "stall samples", tw)
for ln, (c, w, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    f, n = ln
    if f not in src:
        try:
            # NCU_LINES_SRC: the source directory the profiled binary was built from (default: this tree's csrc)
            import os
            sdir = os.environ.get("NCU_LINES_SRC", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sln_amodal_b200", "csrc"))
            src[f] = open(os.path.join(sdir, f)).read().split("\n")
        except Exception:
            src[f] = []
    text = src[f][n - 1].strip()[:90] if 0 < n <= len(src[f]) else ""
    dom = sorted(zip(st, [h[i] for i in stall_cols]), reverse=True)[:2]
    print(f"{100*c/max(tot,1):5.1f}% inst  {100*w/max(tw,1):5.1f}% stall  {f}:{n:<4} {text}   [{', '.join(f'{n_}={v}' for v, n_ in dom if v)}]")
