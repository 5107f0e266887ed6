"""Exploration: error of the fp32 mode against the golden fixtures (run on a GPU box)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pistachio_b200.engine import get_engine
eng = get_engine(0)
z = np.load(os.path.join(ROOT, "tests", "golden", "reference_cases.npz"), allow_pickle=False)
names = sorted({k.split("/")[0] for k in z.files})
POLS = ("s-wave", "p-wave")
def run(name, flags):
    c = {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(name + "/")}
    if "T" not in c or "n" not in c or "theta" not in c: return None
    n = np.asarray(c["n"])
    res = eng.eval_grid(c["wl"], c["theta"], n.real, n.imag, list(range(n.shape[0])), list(c["d"]), POLS, flags=flags, cos_theta=np.cos(c["theta"])).numpy()
    return res, c
worst = {}
for name in names:
    out = run(name, 512)
    if out is None: continue
    res, c = out
    T, R = res["T"][0], res["R"][0]
    ok = np.isfinite(c["T"]) & np.isfinite(c["R"])
    if not ok.any(): continue
    dT, dR = np.abs(T - c["T"])[ok], np.abs(R - c["R"])[ok]
    relT = (dT / np.maximum(np.abs(c["T"][ok]), 1e-3)).max(); relR = (dR / np.maximum(np.abs(c["R"][ok]), 1e-3)).max()
    if not name.startswith("fuzz") or max(dT.max(), dR.max()) > 1e-5:
        print(f"{name:16s} max|dT| {dT.max():.2e} max|dR| {dR.max():.2e} rel(T) {relT:.2e} rel(R) {relR:.2e} L={len(c['d'])} finiteOK={np.isfinite(T[ok]).all()}")
    worst[name] = max(dT.max(), dR.max())
f = [v for k, v in worst.items() if k.startswith("fuzz")]
print("fuzz stacks:", len(f), "median", np.median(f), "p90", np.percentile(f, 90), "max", max(f))
