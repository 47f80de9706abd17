"""Multi-GPU driver for the batched frame path: one process per GPU, ``torch.distributed`` plumbing.

Replaces the reference's only data-parallel construct, ``multiprocessing.Pool.starmap`` over frames
(``src/bosonic_qiskit/animate.py:191-206``), and the sequential frame loops
(``wigner.py:92-96``, ``animate.py:326-352``).  The path shards with no data-path collective:

  * batched frames (configs c3, c4): contiguous frame blocks per rank;
  * one huge grid (config c5): contiguous row slabs per rank, every rank holds the full rho;

and ends with ONE gather of the finished tiles on the destination rank.  Two gather transports:

  ``"nccl"``  ``torch.distributed.gather`` of equal, padded blocks (NCCL over NVLink; gloo on CPU);
  ``"peer"``  the destination rank exports its output buffer through CUDA IPC, every rank's Wigner
              kernel stores its finished tiles straight into that buffer over NVLink (posted
              writes, overlapped with the fp64 math tile by tile), then one barrier.

torch is used for device memory, streams and the process group only.
"""

from __future__ import annotations

import math
import weakref
from typing import Sequence

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block ``[lo, hi)`` of ``n`` units owned by ``rank``; the first ``n % world`` ranks get one more."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank / world size")
    base, extra = divmod(int(n), world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def shard_sizes(n: int, world: int) -> list[int]:
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def _dist():
    import torch.distributed as dist

    return dist


def _world(group=None) -> tuple[int, int]:
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def _gather_blocks(local, sizes: Sequence[int], dst: int, group=None):
    """Gather per-rank blocks (first dimension = units, possibly unequal) on ``dst``; returns the
    concatenation on ``dst`` and ``None`` elsewhere.  Blocks are padded to the largest block so a
    single equal-count collective is used."""
    import torch

    dist = _dist()
    rank, world = _world(group)
    if world == 1:
        return local
    biggest = max(sizes)
    if local.shape[0] < biggest:
        pad = torch.zeros((biggest - local.shape[0], *local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    local = local.contiguous()
    if rank == dst:
        buf = torch.empty((world, *local.shape), dtype=local.dtype, device=local.device)
        dist.gather(local, list(buf.unbind(0)), dst=dst, group=group)
        return torch.cat([buf[r, : sizes[r]] for r in range(world)], dim=0)
    dist.gather(local, None, dst=dst, group=group)
    return None


def frames_to_wigner_sharded(psi, traced: Sequence[int], xvec, g: float = math.sqrt(2), method: str = "clenshaw",
                             engine=None, group=None, dst: int = 0, gather: str | None = "nccl"):
    """Batched frame loop sharded by frame block.

    ``psi``: (F, 2**n) complex128, the same array on every rank (each rank slices its block).
    Returns ``(W, (lo, hi))``: with ``gather`` set, ``W`` is the full (F, S, S) result on ``dst`` and
    ``None`` elsewhere; with ``gather=None`` it is this rank's (hi-lo, S, S) block (results stay sharded).
    CUDA engines return torch CUDA tensors; host engines (tests) return CPU tensors.
    """
    import torch

    from .engine import get_engine

    rank, world = _world(group)
    eng = engine if engine is not None else get_engine()
    psi = np.ascontiguousarray(psi, dtype=np.complex128)
    if psi.ndim != 2:
        raise ValueError("psi must be (frames, 2**n)")
    F = psi.shape[0]
    S = len(xvec)
    lo, hi = shard_range(F, rank, world)
    sizes = shard_sizes(F, world)
    on_gpu = hasattr(eng, "state_to_wigner_t") and getattr(eng, "device", None) is not None

    if on_gpu:
        dev = torch.device("cuda", eng.device)
        x_t = torch.as_tensor(np.asarray(xvec, dtype=np.float64), device=dev)
        psi_t = torch.from_numpy(psi[lo:hi]).to(dev) if hi > lo else torch.empty((0, psi.shape[1]), dtype=torch.complex128, device=dev)
        if gather == "peer" and world > 1:
            return _peer_store_frames(eng, psi_t, traced, x_t, g, method, F, S, lo, hi, rank, world, dst, group), (lo, hi)
        W_local = (eng.state_to_wigner_t(psi_t, traced, x_t, g=g, method=method) if hi > lo
                   else torch.empty((0, S, S), dtype=torch.float64, device=dev))
    else:
        W_np = eng.state_to_wigner(psi[lo:hi], traced, xvec, g=g, method=method) if hi > lo else np.empty((0, S, S))
        W_local = torch.from_numpy(np.ascontiguousarray(W_np))
    if gather is None:
        return W_local, (lo, hi)
    return _gather_blocks(W_local, sizes, dst, group), (lo, hi)


def wigner_row_slabs(rho, xvec, yvec=None, g: float = math.sqrt(2), method: str = "clenshaw", engine=None, group=None,
                     dst: int = 0, gather: str | None = "nccl"):
    """One large grid on several GPUs (config c5).  Mirror-symmetric grids (what the reference builds) are shared out
    by UNIT: every rank evaluates an equal range of the grid's pairs of grid lines with the symmetric-grid kernels
    and one reduce assembles the grid on ``dst`` (``_shared_units_reduce``).  Any other grid is sharded by row slab:
    rank r evaluates rows ``[lo, hi)`` of ``yvec`` for every column and the slabs are gathered.  Every rank needs
    the full ``rho``.  Same return convention as above with the first dimension being grid rows."""
    import torch

    from .engine import get_engine

    rank, world = _world(group)
    eng = engine if engine is not None else get_engine()
    xvec = np.asarray(xvec, dtype=np.float64)
    yvec = xvec if yvec is None else np.asarray(yvec, dtype=np.float64)
    ny = len(yvec)
    lo, hi = shard_range(ny, rank, world)
    sizes = shard_sizes(ny, world)
    on_gpu = hasattr(eng, "wigner_t") and getattr(eng, "device", None) is not None
    if on_gpu:
        dev = getattr(eng, "torch_device", None) or torch.device("cuda", eng.device)
        rho_t = torch.as_tensor(np.ascontiguousarray(rho, dtype=np.complex128), device=dev)
        x_t = torch.as_tensor(xvec, device=dev)
        if world > 1 and gather is not None and (yvec is xvec or (len(yvec) == len(xvec) and np.array_equal(yvec, xvec))):
            taken, shared = _shared_units_reduce(eng, rho_t, x_t, xvec, g, method, rank, world, dst, group)
            if taken:  # (``shared`` is None on every rank but ``dst``)
                return shared, (lo, hi)
        y_t = torch.as_tensor(np.ascontiguousarray(yvec[lo:hi]), device=dev)
        W_local = (eng.wigner_t(rho_t, x_t, y_t, g=g, method=method) if hi > lo
                   else torch.empty((0, len(xvec)), dtype=torch.float64, device=dev))
    else:
        W_np = eng.wigner(rho, xvec, yvec[lo:hi], g=g, method=method) if hi > lo else np.empty((0, len(xvec)))
        W_local = torch.from_numpy(np.ascontiguousarray(W_np))
    if gather is None:
        return W_local, (lo, hi)
    return _gather_blocks(W_local, sizes, dst, group), (lo, hi)


def _shared_units_reduce(eng, rho_t, x_t, xvec, g, method, rank, world, dst, group):
    """One mirror-symmetric grid on several GPUs: every rank evaluates an equal share of the grid's units (pairs of
    grid lines, one Clenshaw recurrence for 8 points) into a zeroed full-size frame; the shares touch disjoint
    points, so one reduce (sum) to ``dst`` assembles the grid exactly.  Returns ``(taken, W)``: ``taken`` is False
    when this grid / cutoff / engine cannot take that path (the caller then shards by row slab) and every rank
    decides alike; ``W`` is the grid on ``dst`` and None elsewhere."""
    import torch

    from ._lib import BQ_ERR_UNSUPPORTED, BQError

    if world < 2 or not getattr(eng, "_grid_sharing", False) or not eng.grid_is_mirror(xvec):
        return False, None
    S = len(xvec)
    total = eng.grid_units(S)
    first, last = shard_range(total, rank, world)
    W = torch.zeros((S, S), dtype=torch.float64, device=rho_t.device)
    # Every rank reaches the agreement all_reduce below whatever happens on it: 1 = share evaluated, 0 = this grid /
    # cutoff cannot take the path (all ranks then shard by row slab), -1 = a real failure (out of memory, a CUDA
    # error, a bad range) -- re-raised here and reported as a matching error on the other ranks instead of leaving
    # them waiting in the collective.
    ok, failure = 1, None
    try:
        eng.wigner_t(rho_t, x_t, g=g, method=method, out=W.unsqueeze(0), unit_range=(first, last - first))
    except BQError as exc:
        if exc.code == BQ_ERR_UNSUPPORTED:
            ok = 0
        else:
            ok, failure = -1, exc
    except Exception as exc:  # noqa: BLE001 - anything else must not desynchronise the ranks either
        ok, failure = -1, exc
    flag = torch.tensor([ok], device=rho_t.device)
    _dist().all_reduce(flag, op=_dist().ReduceOp.MIN, group=group)
    verdict = int(flag.item())
    if failure is not None:
        raise failure
    if verdict < 0:
        raise RuntimeError("a peer rank failed while evaluating its share of the grid (see that rank's error)")
    if verdict == 0:
        return False, None
    _dist().reduce(W, dst=dst, op=_dist().ReduceOp.SUM, group=group)
    return True, (W if rank == dst else None)


# ------------------------------------------------------------------------------------------------
# fused compute + gather: kernels store finished tiles directly into the destination rank's buffer
# ------------------------------------------------------------------------------------------------
class _RawCudaArray:
    """``__cuda_array_interface__`` view of a raw float64 device allocation (lets torch wrap it without a copy)."""

    def __init__(self, ptr: int, shape, owner):
        self._owner = owner
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class PeerGatherBuffer:
    """A (units, ...) float64 buffer on ``dst`` that every rank can address (CUDA IPC).

    Construction is collective.  ``ptr_for(unit)`` gives each rank a raw device pointer into the
    destination buffer at that unit's offset; the Wigner kernel's stores then cross NVLink as
    ordinary global stores.  ``finish()`` is the completion barrier.
    """

    def __init__(self, engine, shape, dst: int = 0, group=None):
        import torch

        dist = _dist()
        self.engine, self.dst, self.group = engine, dst, group
        self.rank, self.world = _world(group)
        self.shape = tuple(int(s) for s in shape)
        self.unit_bytes = int(np.prod(self.shape[1:], dtype=np.int64)) * 8
        self.tensor = None
        self._owned = None
        payload = [None]
        if self.rank == dst:
            nbytes = int(np.prod(self.shape, dtype=np.int64)) * 8
            self._owned = engine.device_alloc(nbytes)  # whole cudaMalloc allocation: IPC-exportable at offset 0
            # safety net: a buffer that is dropped without free() gives its memory back when it is collected
            self._finalizer = weakref.finalize(self, PeerGatherBuffer._release, weakref.ref(engine), self._owned)
            self.tensor = torch.as_tensor(_RawCudaArray(self._owned, self.shape, self), device=torch.device("cuda", engine.device))
            payload = [engine.ipc_export(self._owned)]
        if self.world > 1:
            dist.broadcast_object_list(payload, src=dst, group=group)
        self._opened = None
        if self.rank == dst:
            self.base = self._owned
        else:
            self._opened = engine.ipc_open(payload[0])
            self.base = self._opened

    def ptr_for(self, unit: int) -> int:
        return self.base + unit * self.unit_bytes

    def finish(self):
        import torch

        torch.cuda.synchronize()
        if self.world > 1:
            _dist().barrier(group=self.group)
        return self.tensor

    def close(self):
        """Peers unmap the buffer; the owner frees it (only after every user of ``tensor`` is done)."""
        if self._opened is not None:
            self.engine.ipc_close(self._opened)
            self._opened = None

    def free(self):
        if self._owned is not None:
            self.tensor = None
            self._finalizer.detach()
            self.engine.device_free(self._owned)
            self._owned = None

    @staticmethod
    def _release(engine_ref, ptr):
        eng = engine_ref()
        if eng is not None:
            try:
                eng.device_free(ptr)
            except Exception:  # noqa: BLE001 - interpreter shutdown
                pass


def _peer_store_frames(eng, psi_t, traced, x_t, g, method, F, S, lo, hi, rank, world, dst, group):
    """Frames through a CUDA-IPC gather buffer on ``dst``.  The buffer lives for this call only: the result is handed
    back as a torch-owned copy on ``dst`` and the raw allocation is freed once every peer has unmapped it."""
    buf = PeerGatherBuffer(eng, (F, S, S), dst=dst, group=group)
    result = None
    try:
        if hi > lo:
            eng.state_to_wigner_t(psi_t, traced, x_t, g=g, method=method, out=buf.ptr_for(lo))
        out = buf.finish()
        if rank == dst:
            result = out.clone()
    finally:
        buf.close()
        if world > 1:
            _dist().barrier(group=group)  # every peer has unmapped the buffer before its owner frees it
        buf.free()
    return result


__all__ = ["shard_range", "shard_sizes", "frames_to_wigner_sharded", "wigner_row_slabs", "PeerGatherBuffer"]
