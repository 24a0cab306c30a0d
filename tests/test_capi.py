"""The C-ABI library loads without a GPU and exports every symbol include/trace_cuda.h declares; the ctypes structs
match the C layouts; no compute entry point is called here."""
import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "trace_cuda.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(trace_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(T):
    from trace_jl_b200 import _lib as tl
    lib = tl.load()
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/trace_cuda.h but not exported"
    assert set(syms) == set(tl.SIGNATURES), set(syms) ^ set(tl.SIGNATURES)
    assert lib.trace_abi_version() == 3


def test_struct_layouts(T):
    from trace_jl_b200 import _lib as tl
    assert tl.node_dtype.itemsize == 32 and tl.prim_dtype.itemsize == 16
    assert tl.sphere_dtype.itemsize == 160 and tl.material_dtype.itemsize == 44 and tl.light_dtype.itemsize == 164
    assert C.sizeof(tl.FilmDesc) == 16 + 8 + 1024 + 4
    assert C.sizeof(tl.Camera) == 64 + 64 + 16
    assert C.sizeof(tl.SceneDesc) == 14 * 8
    assert C.sizeof(tl.Stats) == 35 * 8      # 12 scalars + 2 x 9 per-class + 5 counters (include/trace_cuda.h trace_stats)


def test_create_fails_loudly_without_gpu(T):
    """No CPU fallback: without a CUDA device trace_create must fail (on a GPU box it succeeds)."""
    from trace_jl_b200 import _lib as tl
    lib = tl.load()
    h = C.c_void_p()
    rc = lib.trace_create(C.byref(h), 0, None)
    import torch
    if torch.cuda.is_available():
        assert rc == 0
        lib.trace_destroy(h)
    else:
        assert rc != 0 and not h.value
        try:
            T.Context(0)
            assert False, "Context() must raise without a GPU"
        except RuntimeError as e:
            assert "no CPU fallback" in str(e)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "trace.jl_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "trace_ref" not in txt and "oracle_lib" not in txt and "oracle/" not in txt, f
