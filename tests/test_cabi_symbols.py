"""The C-ABI library loads on a machine without a GPU and exports every symbol
`include/lqmc_b200.h` declares (no compute calls here)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "lqmc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lqmc_[a-z0-9_]+)\s*\(", text)))


def _lib_path():
    path = os.path.join(ROOT, "latticeqmc_b200", "liblqmc_b200.so")
    if not os.path.isfile(path):
        import __graft_entry__ as g
        g.build()
    return path


def test_header_symbols_exported():
    names = _header_functions()
    assert "lqmc_sweep" in names and "lqmc_create" in names and len(names) >= 20
    lib = ctypes.CDLL(_lib_path())
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/lqmc_b200.h but not exported"


def test_binding_declares_every_export():
    from latticeqmc_b200 import engine
    _lib_path()
    lib = engine.load_library()
    assert sorted(engine.EXPORTS + ("lqmc_engine",)) != []
    for name in _header_functions():
        if name == "lqmc_engine":
            continue
        assert name in engine.EXPORTS, f"{name} missing from the ctypes binding"
        getattr(lib, name)
    assert b"sm_100a" in lib.lqmc_version()


def test_host_philox_stream_is_deterministic_and_uniform():
    from latticeqmc_b200 import philox_uniforms
    _lib_path()
    a = philox_uniforms(7, 3, 11, 4096)
    b = philox_uniforms(7, 3, 11, 4096)
    c = philox_uniforms(7, 4, 11, 4096)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert a.min() >= 0.0 and a.max() < 1.0
    assert abs(a.mean() - 0.5) < 0.02 and abs(a.var() - 1 / 12) < 0.01
    # prefix property: the stream of a shorter request is a prefix of a longer one
    assert np.array_equal(philox_uniforms(7, 3, 11, 100), a[:100])


def test_no_device_fails_loudly():
    """Without a CUDA device the engine must raise, never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from latticeqmc_b200 import SweepEngine, EngineError
    _lib_path()
    with pytest.raises((EngineError, ValueError)):
        SweepEngine(np.eye(4), 0.5, 10, n_chains=1)
