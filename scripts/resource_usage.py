"""Static resource evidence: per kernel of libislands_b200.so, registers per thread, static shared memory, stack
(local-memory frame) and whether the SASS holds local loads / stores (spills), from `cuobjdump --dump-resource-usage`
and `cuobjdump -sass`.  Usage: python scripts/resource_usage.py > profiles/r02_resource_usage.txt"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "islands_b200/lib/libislands_b200.so"
run = lambda *a: subprocess.run(a, capture_output=True, text=True).stdout
res = run("cuobjdump", "--dump-resource-usage", LIB)
usage = collections.OrderedDict()
cur = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
    if m and cur:
        usage[cur] = tuple(int(x) for x in m.groups())
        cur = None
local_ops = collections.Counter()
cur = None
for line in run("cuobjdump", "-sass", LIB).splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
    if m and cur and (m.group(1).startswith("LDL") or m.group(1).startswith("STL")):
        local_ops[cur] += 1
names = run("cu++filt", *usage.keys()).splitlines() if usage else []
print(f"# {LIB}: per-kernel resources (cuobjdump --dump-resource-usage; LDL/STL counted in the SASS), sm_100a")
print("# dynamic shared memory (row ring, tables, result arrays) is set per launch and is not part of SHARED")
print("kernel | registers | static shared B | stack B | LDL+STL in SASS")
for (fn, (reg, stack, shared, local)), name in zip(usage.items(), names):
    name = re.sub(r"\(isl::SearchArgs\)", "", name)[:120]
    print(f"{name} | {reg} | {shared} | {stack} | {local_ops[fn]}")
print(f"# kernels: {len(usage)}; with a stack frame: {sum(1 for v in usage.values() if v[1])}; "
      f"with local loads/stores: {sum(1 for f in usage if local_ops[f])}")
