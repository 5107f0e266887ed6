"""Multi-GPU sharding of a sweep: one process per GPU, no data-path collective.

Every spectral point is independent, so the grid is cut into contiguous slabs along its
slowest axis (the thickness sweep if there is one, otherwise wavelength) and each rank evaluates its slab with the
full (tiny) material/layer tables.  Results stay sharded unless the caller asks for them on every device, in which
case one all-gather (NCCL over NVLink on GPUs; gloo in the CPU tests) assembles [sweep][pol][wl][theta] arrays.
Reference seam: Structure.angle_resolved_spectra's multiprocessing fan-out (transfer_matrix.py:597-628).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous, balanced [lo, hi) of `n` items for `rank`; the first n % world ranks get one extra item."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_axis(n_sweep: int, n_wl: int) -> str:
    return "sweep" if n_sweep > 1 else "wl"


def init_from_env(backend: str | None = None):
    """(rank, world, local_rank) from torchrun's environment; initialises torch.distributed when world > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_prepared(prep: dict, world: int, rank: int) -> dict:
    """Slice the tensors of `TMMEngine.prepare_workload` to this rank's slab (views, no copies of tables)."""
    out = dict(prep)
    n_sweep = 1 if prep["sweep_d"] is None else prep["sweep_d"].numel()
    n_wl = prep["wl"].numel()
    axis = shard_axis(n_sweep, n_wl)
    if axis == "sweep":
        lo, hi = shard_bounds(n_sweep, world, rank)
        out["sweep_d"] = prep["sweep_d"][lo:hi].contiguous()
    else:
        lo, hi = shard_bounds(n_wl, world, rank)
        out["wl"] = prep["wl"][lo:hi].contiguous()
        out["mat_n"] = prep["mat_n"][:, lo:hi].contiguous()
        out["mat_k"] = prep["mat_k"][:, lo:hi].contiguous()
    out["shard"] = (axis, lo, hi)
    return out


def allgather_result(local: torch.Tensor, axis: str, sizes, group=None) -> torch.Tensor:
    """Assemble a [sweep][pol][wl][theta] tensor from per-rank slabs cut along `axis` ('sweep' -> dim 0,
    'wl' -> dim 2).  `sizes` lists every rank's extent along that axis.  Equal slabs use one
    all_gather_into_tensor; ragged slabs fall back to all_gather on padded buffers."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    if axis == "sweep" and len(set(sizes)) == 1:
        # slabs along dim 0 concatenate directly: one collective, no staging copies
        full = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(full, local.contiguous(), group=group)
        return full
    if axis == "wl" and len(set(sizes)) == 1:
        # every (sweep, pol) plane [wl][theta] is contiguous in both the slab and the assembled tensor
        n_sw, n_pol, n_wl, n_th = local.shape
        full = torch.empty((n_sw, n_pol, world * n_wl, n_th), dtype=local.dtype, device=local.device)
        local = local.contiguous()
        for i in range(n_sw):
            for j in range(n_pol):
                dist.all_gather_into_tensor(full[i, j], local[i, j], group=group)
        return full
    dim = 0 if axis == "sweep" else 2
    moved = local.movedim(dim, 0).contiguous()  # ragged slabs: pad to the largest, gather, trim
    cap = max(sizes)
    pad = torch.zeros((cap,) + tuple(moved.shape[1:]), dtype=moved.dtype, device=moved.device)
    pad[: moved.shape[0]] = moved
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    full = torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)
    return full.movedim(0, dim).contiguous()


class AssembledResult:
    """R, T, A of a sweep sharded along its slowest axis, ASSEMBLED ON EVERY GPU by the compute kernel itself.

    Each array is a symmetric-memory allocation (torch.distributed._symmetric_memory: same layout on every rank,
    peer-mapped, bound to one NVLink multicast object).  The compute kernel writes every result straight into all
    GPUs' copies, so the transfer overlaps the arithmetic store by store and no all-gather pass follows:

      mode "unicast"    one store per receiving GPU through the peer-mapped addresses (pst_peer_outputs);
                        per-GPU link traffic (N-1)/N of the array in each direction
      mode "multicast"  one multimem.st per result (PST_FLAG_MULTICAST_OUT); the NVSwitch replicates it -- the sender's own
                        copy comes back through the switch as well, so every GPU receives the WHOLE array
      mode "nccl"       compute into the local slab, then the in-place NCCL all-gather (two passes; the baseline)

    `root` = None assembles on EVERY GPU; `root` = r assembles on rank r only (north_star: "a final gather only when
    results must be assembled on one device"): every rank stores its slab into rank r's copy and nowhere else, so one
    link receives and the other GPUs' links and memory stay free for their next chunk (NCCL mode: dist.gather).
    `keys` selects the arrays that cross the links -- ("T", "R") by default: A = (1 - R) - T exactly (it is computed
    that way, pst_chain.cuh finish_point_det), so whoever needs A recomputes it from the assembled T and R
    (`absorptance()`), and a third of the wire bytes never travel.

    `finish()` orders the devices (stream sync + symmetric-memory barrier) before anyone reads `full[...]`.
    """

    def __init__(self, n_slow: int, plane_shape, rank: int, world: int, device, group=None, mode: str = "unicast",
                 root: int | None = None, keys=("T", "R")):
        import torch.distributed._symmetric_memory as symm

        if n_slow % world:
            raise ValueError("AssembledResult needs equal slabs (slow axis divisible by the world size)")
        if mode not in ("unicast", "multicast", "nccl"):
            raise ValueError("mode must be 'unicast', 'multicast' or 'nccl'")
        if root is not None and not (0 <= root < world):
            raise ValueError("root must be a rank of the group")
        if mode == "multicast" and root is not None:
            raise ValueError("multicast stores reach every GPU of the mapping: use mode='unicast' with a root")
        self.rank, self.world, self.per = rank, world, n_slow // world
        self.root, self.keys = root, tuple(keys)
        self.group = group if group is not None else dist.group.WORLD
        self.shape = (n_slow,) + tuple(plane_shape)
        self.plane_elems = 1
        for s in plane_shape:
            self.plane_elems *= int(s)
        self.full, self.handles = {}, {}
        for k in self.keys:
            t = symm.empty(self.shape, dtype=torch.float64, device=device)
            self.full[k] = t
            self.handles[k] = symm.rendezvous(t, self.group)
        if mode == "multicast" and not all(int(h.multicast_ptr) != 0 for h in self.handles.values()):
            mode = "unicast"  # fabric without multicast support
        self.mode = mode

    @property
    def _slab_offset(self) -> int:
        return self.rank * self.per * self.plane_elems * 8

    def local_slab(self, key: str) -> torch.Tensor:
        lo = self.rank * self.per
        return self.full[key][lo:lo + self.per]

    def multicast_ptrs(self) -> dict:
        return {k: int(h.multicast_ptr) + self._slab_offset for k, h in self.handles.items()}

    def peer_ptrs(self) -> dict:
        """This rank's slab inside every GPU's copy, rotated so that each rank starts with its right-hand neighbour
        (at any instant the ranks then target different destinations)."""
        order = [(self.rank + 1 + i) % self.world for i in range(self.world)] if self.root is None else [self.root]
        return {k: [int(h.buffer_ptrs[i]) + self._slab_offset for i in order] for k, h in self.handles.items()}

    def absorptance(self) -> torch.Tensor:
        """A = (1 - R) - T of the assembled arrays: bit for bit what the kernels compute per point."""
        return (1.0 - self.full["R"]) - self.full["T"]

    def eval_into(self, eng, prep_slab: dict, pols, flags: int = 0):
        """Evaluate this rank's slab (a prepared dict whose sweep_d holds this rank's `per` thicknesses)."""
        from ._cabi import FLAG_MULTICAST_OUT

        if self.mode == "multicast":
            return eng.eval_prepared(prep_slab, pols=pols, out_ptrs=self.multicast_ptrs(), flags=flags | FLAG_MULTICAST_OUT)
        if self.mode == "unicast":
            return eng.eval_prepared(prep_slab, pols=pols, out_ptrs=self.peer_ptrs(), flags=flags)
        # the slab as the engine's [sweep][pol][wl][theta] view (a slab cut along the wavelength axis carries a leading 1)
        def as_grid(t):
            return t.view((-1,) + tuple(t.shape[-3:]))

        return eng.eval_prepared(prep_slab, pols=pols, want=self.keys, out={k: as_grid(self.local_slab(k)) for k in self.keys},
                                 flags=flags)

    def finish(self):
        if self.mode == "nccl":
            for k in self.keys:
                if self.root is None:
                    dist.all_gather_into_tensor(self.full[k], self.local_slab(k), group=self.group)
                else:
                    slabs = [self.full[k][r * self.per:(r + 1) * self.per] for r in range(self.world)] if self.rank == self.root else None
                    dist.gather(self.local_slab(k), gather_list=slabs, dst=self.root, group=self.group)
            return
        torch.cuda.current_stream().synchronize()
        self.handles[self.keys[0]].barrier()
