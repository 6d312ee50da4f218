"""Drop-in ``envs`` package: put this directory in place of (or ahead of) the reference's ``src/envs`` and
``train_quadruped.py`` / ``eval_quadruped.py`` run unchanged on the B200 path ("scripts unchanged").
Every module re-exports the classes of ``quadruped_gym_b200`` under the reference's module and class names."""
