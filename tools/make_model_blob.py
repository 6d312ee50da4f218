"""Compile the reference MJCF into the packaged model blob.

    python tools/make_model_blob.py [/root/reference/src/models/quadruped/scene.xml]

Writes quadruped_gym_b200/model/assets/mg996r_scene.qgblob (numbers only; the MJCF/OBJ sources stay
in the reference checkout).
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from quadruped_gym_b200.model import DEFAULT_BLOB, compile_mjcf  # noqa: E402

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/models/quadruped/scene.xml"
cm = compile_mjcf(src)
os.makedirs(os.path.dirname(DEFAULT_BLOB), exist_ok=True)
with open(DEFAULT_BLOB, "wb") as fh:
    fh.write(cm.to_blob())
print(f"wrote {DEFAULT_BLOB} ({os.path.getsize(DEFAULT_BLOB)} bytes) from {src}")
