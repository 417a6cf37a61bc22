import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def uniform(rng, n, d):
    """Reference test/bench distribution: uniform [-1, 1) f32 (leann.rs:1078-1083, benches/*:9-14)."""
    return (rng.rand(n, d).astype(np.float32) * 2 - 1).astype(np.float32)


@pytest.fixture(scope="session")
def orc():
    from oracle import pyoracle

    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def gpu_lib():
    from islands_b200 import _ffi

    lib = _ffi.load()
    if lib.isl_device_count() < 1:
        pytest.fail("GPU test selected but no CUDA device is usable (no CPU fallback exists)")
    return lib


_GRAPH_CACHE = {}


def oracle_graph(orc, n, d, seed=0, metric=0, batch=16, dup=0, **cfg_kw):
    """Oracle-built LEANN graph on uniform data (cached per session). dup > 0 appends copies of
    the first `dup` vectors so that exact distance ties occur."""
    from islands_b200 import LeannConfig

    key = (n, d, seed, metric, batch, dup, tuple(sorted(cfg_kw.items())))
    if key not in _GRAPH_CACHE:
        rng = np.random.RandomState(seed)
        v = uniform(rng, n, d)
        if dup:
            v[n - dup:] = v[:dup]
        cfg = LeannConfig(metric=metric, **cfg_kw)
        levels = orc.draw_levels(seed + 1, n, cfg.ml, cfg.max_layers)
        off, nbrs, entry, _ = orc.leann_build(cfg._s, v, levels, batch=batch, threads=os.cpu_count() or 1)
        _GRAPH_CACHE[key] = (cfg, v, levels, off, nbrs, entry)
    return _GRAPH_CACHE[key]
