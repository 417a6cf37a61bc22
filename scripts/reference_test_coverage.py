"""Which of the reference's own unit tests (the `#[cfg(test)]` modules of src/core/{distance,hnsw,leann,pq,search,
storage}.rs) does this repository's suite re-express?  A reference test counts as covered when a file under tests/
cites a line range of that source file that overlaps the test's body.  Needs /root/reference (it is not present on the
GPU box): run here; the result is committed as tests/golden/reference_test_coverage.json and a CPU test checks that
every citation it lists is still in the suite.
Usage: python scripts/reference_test_coverage.py [--write]"""
import collections
import glob
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src/core/"
FILES = ["distance", "hnsw", "leann", "pq", "search", "storage"]


def reference_tests():
    out = {}
    for f in FILES:
        lines = open(REF + f + ".rs").read().split("\n")
        start = next(i for i, l in enumerate(lines) if "mod tests" in l)
        items = [(i + 1, m.group(1)) for i, l in enumerate(lines) if i > start
                 for m in [re.match(r"\s+fn (test_\w+|prop_\w+)\(", l)] if m]
        out[f] = [(a, name, items[j + 1][0] - 1 if j + 1 < len(items) else len(lines)) for j, (a, name) in enumerate(items)]
    return out


def citations():
    cites = collections.defaultdict(list)
    paths = glob.glob(os.path.join(ROOT, "tests", "**", "*.py"), recursive=True) + glob.glob(os.path.join(ROOT, "tests", "cpp", "*.cpp"))
    for p in sorted(paths):
        for m in re.finditer(r"(distance|hnsw|leann|pq|search|storage)\.rs:(\d+)(?:-(\d+))?", open(p).read()):
            a = int(m.group(2))
            cites[m.group(1)].append((a, int(m.group(3) or a), os.path.relpath(p, ROOT)))
    return cites


def main():
    tests, cites = reference_tests(), citations()
    table, missing = {}, []
    for f, ts in tests.items():
        for a, name, e in ts:
            hit = sorted({c[2] for c in cites[f] if not (c[1] < a or c[0] > e)})
            table[f"{f}.rs:{a} {name}"] = hit
            if not hit:
                missing.append(f"{f}.rs:{a} {name}")
    print(f"{len(table)} reference tests, {len(table) - len(missing)} re-expressed, {len(missing)} not cited")
    for m in missing:
        print("  missing:", m)
    if "--write" in sys.argv:
        with open(os.path.join(ROOT, "tests", "golden", "reference_test_coverage.json"), "w") as fh:
            json.dump(table, fh, indent=1, sort_keys=True)
            fh.write("\n")
    return 1 if missing else 0


if __name__ == "__main__":
    sys.exit(main())
