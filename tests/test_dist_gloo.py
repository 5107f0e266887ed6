"""world_size-2 `gloo` tests (CPU) of the multi-GPU host logic: slab bounds, slicing of a prepared workload and the
assembly of [sweep][pol][wl][theta] results.  The per-rank compute is stood in for by the oracle's closed form (the
CUDA path cannot run here); on GPUs the same code runs over NCCL (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, axis, ragged, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from oracle import tmm_oracle as orc
    from pistachio_b200 import dist as pdist

    r, w, _ = pdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    n_wl = 23 if ragged else 24
    n_sw = 5 if ragged else 4
    wl = np.linspace(0.8, 1.2, n_wl)
    th = np.array([0.0, 0.4, 0.9])
    sweep = np.linspace(0.1, 0.5, n_sw)
    n_list = [np.full(n_wl, 1.0 + 0j), np.full(n_wl, 2.3 + 0.01j), np.full(n_wl, 1.4 + 0j), np.full(n_wl, 1.5 + 0j)]

    def compute(wl_idx, sw_idx):
        out = np.empty((len(sw_idx), 2, len(wl_idx), len(th)))
        for i, k in enumerate(sw_idx):
            st = orc.Stack(wl[wl_idx], [n[wl_idx] for n in n_list], [0.0, sweep[k], 0.3, 0.0])
            T, _ = orc.spectrum_closed_form(st, th)
            out[i] = T
        return out

    full_ref = compute(np.arange(n_wl), np.arange(n_sw))
    prep = {"wl": torch.from_numpy(wl), "mat_n": torch.zeros(4, n_wl), "mat_k": torch.zeros(4, n_wl),
            "sweep_d": torch.from_numpy(sweep) if axis == "sweep" else None}
    shard = pdist.shard_prepared(prep, world, rank)
    ax, lo, hi = shard["shard"]
    assert ax == axis
    if axis == "sweep":
        assert shard["sweep_d"].numel() == hi - lo and shard["wl"].numel() == n_wl
        local = compute(np.arange(n_wl), np.arange(lo, hi))
        sizes = [pdist.shard_bounds(n_sw, world, q)[1] - pdist.shard_bounds(n_sw, world, q)[0] for q in range(world)]
    else:
        assert shard["wl"].numel() == hi - lo and shard["mat_n"].shape == (4, hi - lo)
        local = compute(np.arange(lo, hi), np.arange(n_sw))[:1] if False else compute(np.arange(lo, hi), np.arange(n_sw))
        sizes = [pdist.shard_bounds(n_wl, world, q)[1] - pdist.shard_bounds(n_wl, world, q)[0] for q in range(world)]
    got = pdist.allgather_result(torch.from_numpy(local), axis, sizes).numpy()
    ok = got.shape == full_ref.shape and np.array_equal(got, full_ref)
    # timing contract of bench.py: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = ok and t.item() == world
    ret[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("axis,ragged", [("wl", False), ("sweep", False), ("wl", True), ("sweep", True)])
def test_two_rank_shard_and_gather(axis, ragged):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, axis, ragged, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0, f"rank exited with {p.exitcode}"
    assert dict(ret) == {0: True, 1: True}


def test_shard_bounds_cover_everything():
    from pistachio_b200.dist import shard_bounds

    for n in (1, 7, 8, 8192, 100000):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)
