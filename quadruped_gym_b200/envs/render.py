"""Rendering bridge: the batched physics never renders; ``render()`` of a single-environment object copies
that environment's ``qpos`` into a CPU ``mujoco.MjData`` and lets MuJoCo's own renderer draw it, with the frame pacing,
window / video behaviour and overlay geoms of the reference (/root/reference/src/envs/quadruped.py:77-86,184-316,
walking_quad.py:77-86): a frame is produced only when simulated time has advanced by 1/render_fps; "human" mode waits
for the wall clock to catch up with simulated time and shows the frame with OpenCV; "rgb_array" returns it;
``save_video`` appends every produced frame to an mp4.

Needs the ``mujoco`` wheel (and ``cv2`` for windows / video), which this build image does not have: everything here is
imported lazily and fails with a clear error at the first ``render()`` call, never at construction, so that scripts
which build a rendering env but are run headless still work up to that point.  UNTESTED against a real MuJoCo in this
image for that reason; tests/test_dropin.py drives it against a recording stub of the few MuJoCo calls it makes.
"""
from __future__ import annotations

import os
import time
from typing import Optional, Sequence

import numpy as np

DEFAULT_SCENE = "./models/quadruped/scene.xml"      # the reference's default model_path (quadruped.py:41), CWD = src/


class MujocoRenderBridge:
    def __init__(self, model_path: Optional[str], mode: Optional[str], width=720, height=480, fps=30,
                 save_video=False, video_path="videos/simulation.mp4"):
        if mode not in (None, "human", "rgb_array"):
            raise ValueError(f"unknown render_mode {mode!r}")
        self.model_path = model_path or DEFAULT_SCENE
        self.mode, self.width, self.height, self.fps = mode, int(width), int(height), int(fps)
        self.save_video, self.video_path = bool(save_video), video_path
        self._mj = None          # (mujoco module, MjModel, MjData, Renderer, camera, scene option)
        self._writer = None
        self.restart()

    def restart(self):
        """reset(): frame counter and wall-clock origin start over (quadruped.py:126-131)."""
        self._frames = 0
        self._t0 = time.time() if self.mode == "human" else None

    def _lazy(self):
        if self._mj is not None:
            return self._mj
        try:
            import mujoco
        except ImportError as e:
            raise NotImplementedError("render() needs the `mujoco` wheel: the renderer is MuJoCo's own (CPU / OpenGL); "
                                      "the batched B200 physics does not depend on it") from e
        if not os.path.exists(self.model_path):
            raise FileNotFoundError(f"render() needs the MJCF scene to draw: {self.model_path} not found")
        m = mujoco.MjModel.from_xml_path(self.model_path)
        d = mujoco.MjData(m)
        cam = mujoco.MjvCamera()
        cam.distance, cam.elevation, cam.azimuth = 1.0, -30, 120
        opt = mujoco.MjvOption()
        opt.flags[mujoco.mjtVisFlag.mjVIS_JOINT] = False
        opt.flags[mujoco.mjtVisFlag.mjVIS_CONTACTPOINT] = False
        opt.frame = mujoco.mjtFrame.mjFRAME_SITE
        opt.geomgroup[:] = 1
        self._mj = (mujoco, m, d, mujoco.Renderer(m, height=self.height, width=self.width), cam, opt)
        return self._mj

    def _overlay(self, mujoco, scene, item):
        if scene.ngeom >= scene.maxgeom:
            return
        g = scene.geoms[scene.ngeom]
        kind = item[0]
        if kind == "arrow":          # ("arrow", origin, vector, rgba, z offset): 0.2 m per unit, 5 mm shaft
            _, origin, vec, rgba, dz = item
            a = np.asarray(origin, dtype=np.float64) + np.array([0.0, 0.0, dz])
            b = a + 0.2 * np.asarray(vec, dtype=np.float64)
            mujoco.mjv_initGeom(g, mujoco.mjtGeom.mjGEOM_ARROW1, np.zeros(3), np.zeros(3), np.zeros(9), np.asarray(rgba, dtype=np.float32))
            mujoco.mjv_connector(g, mujoco.mjtGeom.mjGEOM_ARROW1, 0.005, a, b)
        else:                        # ("point", position, rgba, _): 1 cm sphere
            _, pos, rgba, _ = item
            mujoco.mjv_initGeom(g, mujoco.mjtGeom.mjGEOM_SPHERE, np.full(3, 0.01), np.asarray(pos, dtype=np.float64),
                                np.eye(3).reshape(9), np.asarray(rgba, dtype=np.float32))
        scene.ngeom += 1

    def frame(self, qpos: np.ndarray, sim_time: float, overlays: Sequence[tuple] = ()):
        """One ``render()`` call of the reference: returns an RGB array in "rgb_array" mode when a frame is due, else None."""
        if self.mode is None and not self.save_video:
            return None
        if self._frames >= int(sim_time * self.fps):
            return None
        self._frames += 1
        mujoco, m, d, renderer, cam, opt = self._lazy()
        d.qpos[:] = np.asarray(qpos, dtype=np.float64)
        d.time = sim_time
        mujoco.mj_forward(m, d)
        cam.lookat[:] = d.qpos[:3]
        renderer.update_scene(d, scene_option=opt, camera=cam)
        for item in overlays:
            self._overlay(mujoco, renderer.scene, item)
        pixels = renderer.render()
        if self.save_video or self.mode == "human":
            import cv2
            bgr = cv2.cvtColor(pixels, cv2.COLOR_RGB2BGR)
            if self.save_video:
                if self._writer is None:
                    os.makedirs(os.path.dirname(self.video_path) or ".", exist_ok=True)
                    self._writer = cv2.VideoWriter(self.video_path, cv2.VideoWriter_fourcc(*"mp4v"), self.fps, (self.width, self.height))
                self._writer.write(bgr)
            if self.mode == "human":
                if self._t0 is None:
                    self._t0 = time.time()
                wait = self._t0 + sim_time - time.time()
                if wait > 0:
                    time.sleep(wait)
                cv2.imshow("Simulation", bgr)
                cv2.waitKey(1)
                return None
        return pixels if self.mode == "rgb_array" else None

    def close(self):
        if self._mj is not None:
            self._mj[3].close()
            self._mj = None
        if self._writer is not None:
            self._writer.release()
            self._writer = None
        if self.mode == "human":
            try:
                import cv2
                cv2.destroyAllWindows()
            except Exception:
                pass
