"""Exploration: error of the fp32 mode against the fp64 default as a function of the stack depth (run on a GPU box with a
library built with -DPST_FP32_MAX_LAYERS=1000000)."""
import numpy as np

from pistachio_b200.engine import get_engine

eng = get_engine(0)
FP32 = 1 << 9
wl = np.linspace(0.9, 1.1, 256)
th = [0.0, 0.4]
for name, nh, nl in (("low contrast 1.46/1.45", 1.46, 1.45), ("ZnS/MgF2 2.32/1.36", 2.32, 1.36)):
    nn = np.array([1.0, nh, nl, 1.5])[:, None] * np.ones((1, wl.size))
    for pairs in (10, 20, 44, 128, 500, 2000, 5000):
        mats = [0] + [1, 2] * pairs + [3]
        d = [0.0] + [1.0 / (4 * nh), 1.0 / (4 * nl)] * pairs + [0.0]
        args = (wl, th, nn, 0 * nn, mats, d, ("s-wave", "p-wave"))
        a, b = eng.eval_grid(*args, flags=FP32).numpy(), eng.eval_grid(*args).numpy()
        print(f"{name:24s} {2 * pairs:6d} layers: max|dR| {np.abs(a['R'] - b['R']).max():.2e}  max|dT| {np.abs(a['T'] - b['T']).max():.2e}")
# chirped (C4-like): no two layers alike
rng = np.random.default_rng(0)
for L in (100, 1000, 10000):
    lam0 = np.linspace(0.8, 1.25, L)
    idx = np.where(np.arange(L) % 2 == 0, 1.46, 1.45)
    nn = np.array([1.0, 1.46, 1.45, 1.5])[:, None] * np.ones((1, wl.size))
    mats = [0] + [1 if i % 2 == 0 else 2 for i in range(L)] + [3]
    d = [0.0] + list(lam0 / (4 * idx)) + [0.0]
    args = (wl, [0.0], nn, 0 * nn, mats, d, ("s-wave",))
    a, b = eng.eval_grid(*args, flags=FP32).numpy(), eng.eval_grid(*args).numpy()
    print(f"chirped 1.46/1.45        {L:6d} layers: max|dR| {np.abs(a['R'] - b['R']).max():.2e}  max|dT| {np.abs(a['T'] - b['T']).max():.2e}")
