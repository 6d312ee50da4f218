"""``from envs.math_utils import exp_dist, unit`` (/root/reference/src/envs/math_utils.py:4-8).  The online
frequency / amplitude estimator of that module runs inside the walking kernel (csrc/qg_walk.cuh)."""
import numpy as np


def exp_dist(x):
    return np.exp(x) - 1


def unit(v):
    return v / np.linalg.norm(v)
