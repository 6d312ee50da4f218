"""Compiled-model blob: the boundary object shared by the CUDA library and the CPU oracle.

The reference builds its model with ``mujoco.MjModel.from_xml_path``
(/root/reference/src/envs/quadruped.py:59).  Here the equivalent product is a flat little-endian
buffer of named sections (TLV) that ``qg_model_load`` (include/quadgym.h) and
``qgo_model_load`` (oracle/qg_oracle.c) both parse.  Two producers can write it: the in-repo MJCF
compiler (``mjcf.py``) and, where ``mujoco`` imports, ``export_mujoco.py``.

Layout::

    magic   8 bytes  b"QGBLOB01"
    nsec    u32, pad u32
    section*:
        name   16 bytes, NUL padded ASCII
        dtype  u32   (1 = float64, 2 = int32)
        count  u32   number of elements
        data   count * itemsize bytes, padded with zeros to a multiple of 8
"""
from __future__ import annotations

import struct
from typing import Dict

import numpy as np

MAGIC = b"QGBLOB01"
_DT_F64 = 1
_DT_I32 = 2


def pack(sections: Dict[str, np.ndarray]) -> bytes:
    out = [MAGIC, struct.pack("<II", len(sections), 0)]
    for name, arr in sections.items():
        if len(name) > 15:
            raise ValueError(f"section name too long: {name}")
        a = np.asarray(arr)
        if a.dtype.kind == "f":
            a = np.ascontiguousarray(a, dtype="<f8").ravel()
            code = _DT_F64
        elif a.dtype.kind in "iub":
            a = np.ascontiguousarray(a, dtype="<i4").ravel()
            code = _DT_I32
        else:
            raise TypeError(f"section {name}: unsupported dtype {a.dtype}")
        raw = a.tobytes()
        pad = (-len(raw)) % 8
        out.append(name.encode("ascii").ljust(16, b"\0"))
        out.append(struct.pack("<II", code, a.size))
        out.append(raw + b"\0" * pad)
    return b"".join(out)


def unpack(buf: bytes) -> Dict[str, np.ndarray]:
    if buf[:8] != MAGIC:
        raise ValueError("not a quadgym model blob (bad magic)")
    nsec, _ = struct.unpack_from("<II", buf, 8)
    off = 16
    out: Dict[str, np.ndarray] = {}
    for _ in range(nsec):
        name = buf[off:off + 16].rstrip(b"\0").decode("ascii")
        code, count = struct.unpack_from("<II", buf, off + 16)
        off += 24
        if code == _DT_F64:
            nbytes = 8 * count
            arr = np.frombuffer(buf, dtype="<f8", count=count, offset=off).copy()
        elif code == _DT_I32:
            nbytes = 4 * count
            arr = np.frombuffer(buf, dtype="<i4", count=count, offset=off).copy()
        else:
            raise ValueError(f"section {name}: bad dtype code {code}")
        off += nbytes + ((-nbytes) % 8)
        out[name] = arr
    return out
