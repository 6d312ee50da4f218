"""The C-ABI shared library loads, exports every symbol include/quadgym.h declares, parses model blobs
and reports errors through return codes (no compute calls here: no GPU in this tier)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from quadruped_gym_b200 import _lib
from quadruped_gym_b200.model import blob as qblob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "quadgym.h")).read()
    declared = set(re.findall(r"\b(qg_[a-z0-9_]+)\s*\(", header))
    declared -= {"qg_counters"}
    assert declared == set(_lib.SYMBOLS)
    L = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in _lib.lib().qg_version()


def test_model_load_and_info(blob):
    L = _lib.lib()
    h = C.c_void_p()
    assert L.qg_model_load(blob, len(blob), C.byref(h)) == 0
    sizes, ts = (C.c_int * 8)(), C.c_double()
    assert L.qg_model_info(h, sizes, C.byref(ts)) == 0
    assert list(sizes) == [19, 18, 12, 14, 13, 25, 5, 33] and ts.value == 0.002
    L.qg_model_destroy(h)


def test_bad_blobs_are_rejected_with_codes(blob):
    L = _lib.lib()
    h = C.c_void_p()
    assert L.qg_model_load(b"not a blob at all", 17, C.byref(h)) == -2          # QG_EBLOB
    assert b"magic" in L.qg_last_error()
    assert L.qg_model_load(blob[: len(blob) // 2], len(blob) // 2, C.byref(h)) == -2
    A = qblob.unpack(blob)
    A2 = dict(A)
    del A2["geom_mu"]
    raw = qblob.pack(A2)
    assert L.qg_model_load(raw, len(raw), C.byref(h)) == -2 and b"geom_mu" in L.qg_last_error()
    # outside the supported model class -> QG_EMODEL
    A3 = {k: v.copy() for k, v in A.items()}
    A3["jnt_axis"] = A3["jnt_axis"].copy()
    A3["jnt_axis"][3:6] = [1.0, 0.0, 0.0]          # first hinge about x
    raw = qblob.pack(A3)
    assert L.qg_model_load(raw, len(raw), C.byref(h)) == -3 and b"axis" in L.qg_last_error()
    A4 = {k: v.copy() for k, v in A.items()}
    A4["opt_f"][6] = 2.0                            # impratio != 1 with the pyramidal cone
    raw = qblob.pack(A4)
    assert L.qg_model_load(raw, len(raw), C.byref(h)) == -3 and b"impratio" in L.qg_last_error()
    A4["opt_i"][1] = 1                              # elliptic cone with impratio 2: supported
    raw = qblob.pack(A4)
    assert L.qg_model_load(raw, len(raw), C.byref(h)) == 0
    L.qg_model_destroy(h)
    A4["opt_i"][1] = 7
    raw = qblob.pack(A4)
    assert L.qg_model_load(raw, len(raw), C.byref(h)) == -3 and b"cone" in L.qg_last_error()
    with pytest.raises(ValueError):
        _lib.check(-3, "qg_model_load")


def test_no_cpu_fallback(blob):
    """Without a CUDA device the batch constructor fails loudly (QG_ECUDA); with one it succeeds."""
    import torch
    L = _lib.lib()
    h, b = C.c_void_p(), C.c_void_p()
    assert L.qg_model_load(blob, len(blob), C.byref(h)) == 0
    rc = L.qg_batch_create(h, 8, 0, C.byref(b))
    if torch.cuda.is_available():
        assert rc == 0
        L.qg_batch_destroy(b)
    else:
        assert rc == -4 and b"no CPU fallback" in L.qg_last_error()
        from quadruped_gym_b200 import VecQuadrupedEnv
        with pytest.raises(_lib.QuadGymLibraryError):
            VecQuadrupedEnv(4, "cpu")
    assert L.qg_batch_create(h, 0, 0, C.byref(b)) == -1     # QG_EINVAL
    L.qg_model_destroy(h)


def test_product_does_not_import_the_oracle():
    """The product path may not route through oracle/ (test infrastructure)."""
    pkg = os.path.join(ROOT, "quadruped_gym_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), f
                assert "libqgoracle" not in txt and "qg_oracle.h" not in txt and "qgo_step" not in txt, f


def test_missing_extension_fails_loudly(tmp_path):
    """No silent fallback: with the .so absent, loading the library raises (checked in a fresh interpreter)."""
    import subprocess, sys
    code = ("import os, sys; sys.path.insert(0, %r); os.environ['QG_LIB'] = %r\n"
            "from quadruped_gym_b200 import _lib\n"
            "try:\n    _lib.lib()\nexcept _lib.QuadGymLibraryError as e:\n    print('RAISED', 'no CPU fallback' in str(e))\n") % (ROOT, str(tmp_path / "nope.so"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.stdout.strip() == "RAISED True", out.stdout + out.stderr


def test_top_level_exports_resolve_lazily():
    """The package face a reference user switches to: every advertised name resolves (no GPU, no .so call needed)."""
    import quadruped_gym_b200 as q
    for name in q.__all__:
        assert getattr(q, name) is not None
    with pytest.raises(AttributeError):
        getattr(q, "NoSuchEnv")
