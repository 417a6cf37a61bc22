"""Hostile-input sweep of the three `from_bytes` entry points (leann.rs:1064-1066, pq.rs:356-358, hnsw.rs:512-514): valid
bincode images (built by oracle/bincode_oracle.py, the CPU restatement of the layout) are truncated, extended and
corrupted — length fields set to huge values in particular — and handed to the library.  Every call must come back with
a status (ISL_SERIALIZATION for malformed input; on a box without a GPU a well-formed image ends in ISL_CUDA_ERROR when
the upload starts) — no crash, no runaway allocation.  Each batch runs in its own process under an address-space limit,
so a segfault, an abort or a multi-gigabyte allocation is seen as a failed batch.  No GPU needed: the parser is host code.
Usage: python scripts/fuzz_from_bytes.py [--cases 3000] [--seed 0]"""
import argparse
import ctypes as C
import os
import resource
import struct
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def images():
    from islands_b200 import HnswConfig, LeannConfig
    from oracle import bincode_oracle as bo

    rng = np.random.RandomState(1)
    n, d = 40, 8
    deg = rng.randint(1, 6, size=n)
    off = np.concatenate([[0], np.cumsum(deg)])
    nbrs = rng.randint(0, n, size=int(off[-1]))
    levels = rng.randint(0, 3, size=n)
    leann = bo.leann_index(LeannConfig(), off, nbrs, levels, 3, int(levels.max()), d)
    cb = rng.rand(4, 16, 2).astype(np.float32)
    pq = bo.product_quantizer(4, 16, 5, 42, cb, d, 1, True)
    vec = rng.rand(n, d).astype(np.float32)
    hnsw = bo.hnsw_graph(HnswConfig(), vec, levels, lambda i, layer: [(i + 1 + layer) % n, (i + 7) % n], 3, int(levels.max()))
    return {"leann": leann, "pq": pq, "hnsw": hnsw, "vectors": vec}


def mutate(rng, data):
    b = bytearray(data)
    kind = rng.randint(0, 7)
    if kind == 0:
        return bytes(b[:rng.randint(0, len(b))])                      # truncation
    if kind == 1:
        return bytes(b) + bytes(rng.randint(0, 256, size=rng.randint(1, 32)).astype(np.uint8))  # trailing bytes
    if kind == 2:                                                      # a u64 field becomes huge
        at = 8 * rng.randint(0, max(1, len(b) // 8))
        huge = [2 ** 64 - 1, 2 ** 63, 2 ** 40, 2 ** 32, len(b), len(b) // 4][rng.randint(0, 6)]
        b[at:at + 8] = struct.pack("<Q", huge)
        return bytes(b[:len(data)])
    if kind == 3:                                                      # the same at an unaligned offset
        at = rng.randint(0, max(1, len(b) - 8))
        b[at:at + 8] = struct.pack("<Q", [2 ** 64 - 1, 2 ** 48, 2 ** 31][rng.randint(0, 3)])
        return bytes(b)
    if kind == 4:                                                      # random byte flips
        for _ in range(rng.randint(1, 8)):
            b[rng.randint(0, len(b))] = rng.randint(0, 256)
        return bytes(b)
    if kind == 5:                                                      # a zeroed run
        at = rng.randint(0, len(b))
        run = min(rng.randint(1, 64), len(b) - at)
        b[at:at + run] = bytes(run)
        return bytes(b)
    return bytes(rng.randint(0, 256, size=rng.randint(0, 200)).astype(np.uint8))  # noise


def run_batch(seed, cases):
    from islands_b200 import _ffi

    resource.setrlimit(resource.RLIMIT_AS, (8 << 30, 8 << 30))  # a runaway allocation fails instead of swapping
    lib = _ffi.load()
    img = images()
    rng = np.random.RandomState(seed)
    vec = np.ascontiguousarray(img["vectors"])
    statuses = {}
    for _ in range(cases):
        which = ("leann", "pq", "hnsw")[rng.randint(0, 3)]
        data = mutate(rng, img[which]) if rng.rand() > 0.02 else img[which]
        buf = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data or b"\0")
        h = C.c_void_p()
        if which == "leann":
            st = lib.isl_index_from_bytes(buf, len(data), vec.ctypes.data_as(_ffi.f32p), vec.shape[1], C.byref(h))
            if st == 0:
                lib.isl_index_free(h)
        elif which == "pq":
            st = lib.isl_pq_from_bytes(buf, len(data), C.byref(h))
            if st == 0:
                lib.isl_pq_free(h)
        else:
            st = lib.isl_hnsw_from_bytes(buf, len(data), C.byref(h))
            if st == 0:
                lib.isl_hnsw_free(h)
        statuses[(which, st)] = statuses.get((which, st), 0) + 1
    print("BATCH_OK", sorted(statuses.items()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=3000)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--batch", type=int, default=-1)
    a = ap.parse_args()
    if a.batch >= 0:
        run_batch(a.batch, a.cases)
        return 0
    per, failed = 250, []
    for b in range((a.cases + per - 1) // per):
        p = subprocess.run([sys.executable, __file__, "--batch", str(a.seed * 1000 + b), "--cases", str(per)],
                           capture_output=True, text=True, timeout=600)
        ok = p.returncode == 0 and "BATCH_OK" in p.stdout
        print(f"batch {b}: {'ok' if ok else 'FAILED rc=' + str(p.returncode)} {p.stdout.strip()[-300:] if ok else p.stderr.strip()[-400:]}")
        if not ok:
            failed.append(b)
    print(f"{a.cases} cases, {len(failed)} failed batches")
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
