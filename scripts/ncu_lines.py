"""Joins the per-instruction page of an .ncu-rep with the line table of the built library (nvdisasm -g), and prints
executed instructions and stall samples aggregated per source line.
Usage: python scripts/ncu_lines.py REPORT.ncu-rep MANGLED_KERNEL_NAME [lib.so] [top]"""
import csv, io, re, subprocess, sys, collections, os, tempfile

rep, fun = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else "islands_b200/lib/libislands_b200.so"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ci, cs = h.index("Instructions Executed"), h.index("# Samples")
inst = [(r[1].strip(), float(r[ci] or 0), float(r[cs] or 0)) for r in rows[hi + 1:] if len(r) > cs]
# line table: extract the cubin, nvdisasm -g -fun
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cub = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
lines = None
for c in cub:
    out = subprocess.run(["nvdisasm", "-g", c], capture_output=True, text=True).stdout
    key = ".text." + fun
    if key not in out:
        continue
    body = out[out.index("//--------------------- " + key):]
    body = body[:body.index("//---------------------", 30)] if "//---------------------" in body[30:] else body
    cur, lines = ("?", 0), []
    for ln in body.splitlines():
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            lines.append(cur)
    break
if lines is None or len(lines) != len(inst):
    print(f"line table ({None if lines is None else len(lines)}) and report ({len(inst)}) disagree", file=sys.stderr)
    n = min(len(lines or []), len(inst))
else:
    n = len(inst)
agg = collections.defaultdict(lambda: [0.0, 0.0, 0])
for i in range(n):
    a = agg[lines[i]]
    a[0] += inst[i][1]
    a[1] += inst[i][2]
    a[2] += 1
ti, ts = sum(v[0] for v in agg.values()) or 1, sum(v[1] for v in agg.values()) or 1
print(f"total warp instructions {ti:.0f}, stall samples {ts:.0f}, {n} SASS instructions")
src = {}
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    if f not in src:
        try:
            src[f] = open(os.path.join("islands_b200/csrc", f)).read().splitlines()
        except Exception:
            src[f] = []
    text = src[f][l - 1].strip()[:100] if 0 < l <= len(src[f]) else ""
    print(f"{100 * v[1] / ts:5.1f}% samples {100 * v[0] / ti:5.1f}% inst ({v[2]:3d} sass) {f}:{l}: {text}")
