"""Where the fused sweep pass spends its time: T-only evaluation vs the device peak search (one C3 chunk)."""
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pistachio_b200 import workloads as wk
from pistachio_b200.engine import get_engine

eng = get_engine(0)
ch = int(sys.argv[1]) if len(sys.argv) > 1 else 512
w = wk.c3_polariton(n_d=ch, n_wl=4096, n_theta=64)
prep = eng.prepare_workload(w)
pols = ("s-wave", "p-wave")
out = {k: torch.empty((ch, 2, 4096, 64), dtype=torch.float64, device=eng.device) for k in ("R", "T", "A")}
nu = 1.0e4 / prep["wl"]

def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); [fn() for _ in range(n)]; b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

t_rta = timed(lambda: eng.eval_prepared(prep, pols=pols, out=out))
t_t = timed(lambda: eng.eval_prepared(prep, pols=pols, want=("T",), out={"T": out["T"]}))
t_pk = timed(lambda: eng.find_peaks(out["T"], height=0.01, distance=20, max_peaks=32, axis=nu, reverse=True))
t_pk_nod = timed(lambda: eng.find_peaks(out["T"], height=0.01, distance=None, max_peaks=32, axis=nu, reverse=True))
gb = out["T"].numel() * 8 / 1e9
print(json.dumps({"chunk": ch, "rta_ms": t_rta, "t_only_ms": t_t, "find_peaks_ms": t_pk, "find_peaks_GBps": gb / t_pk * 1e3,
                  "find_peaks_no_distance_ms": t_pk_nod}))
