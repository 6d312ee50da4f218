"""Model loading: MJCF (compiled here) or a pre-compiled ``.qgblob``.

``DEFAULT_BLOB`` is the compiled form of the reference robot
(/root/reference/src/models/quadruped/scene.xml) produced by ``tools/make_model_blob.py``; it is a
derived artefact (numbers only), shipped so that GPU boxes without the reference checkout can run.
"""
from __future__ import annotations

import os

from .blob import pack, unpack  # noqa: F401
from .mjcf import CompiledModel, compile_mjcf  # noqa: F401

DEFAULT_BLOB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "mg996r_scene.qgblob")

# sensor block of the reference MJCF (quadruped.xml:174-217): name -> (address, dim)
SENSORS = {
    **{f"{j}_{i}_sensor": (3 * (i - 1) + k, 1) for i in range(1, 5) for k, j in enumerate(("hip", "knee", "ankle"))},
    "body_accel": (12, 3), "body_gyro": (15, 3), "body_pos": (18, 3), "body_linvel": (21, 3),
    "body_xaxis": (24, 3), "body_zaxis": (27, 3), "body_vel": (30, 3),
}


def load_model_blob(model_path: str | None = None, mesh_inertia: str = "legacy") -> bytes:
    """``None`` -> packaged blob; ``*.xml`` -> compile the MJCF; anything else is read as a blob."""
    if model_path is None:
        model_path = DEFAULT_BLOB
    if not os.path.exists(model_path):
        raise FileNotFoundError(f"Model file not found: {model_path}")  # quadruped.py:55-56
    if model_path.endswith(".xml"):
        return compile_mjcf(model_path, mesh_inertia=mesh_inertia).to_blob()
    with open(model_path, "rb") as fh:
        return fh.read()
