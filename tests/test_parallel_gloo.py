"""The N > 1 path on CPU: world_size-2 gloo process group, host engine (oracle-backed stand-in)."""

import os
import sys

import numpy as np
import pytest

from bosonic_qiskit_b200.parallel import shard_range, shard_sizes

HERE = os.path.dirname(os.path.abspath(__file__))


def test_shard_ranges_cover_and_balance():
    for n in (0, 1, 7, 8, 512, 4096, 1001):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist

    import workloads
    from bosonic_qiskit_b200.parallel import frames_to_wigner_sharded, wigner_row_slabs
    from oracle_engine import OracleEngine

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = OracleEngine()
    wl = workloads.config3(frames=5, steps=24)  # 5 frames over 2 ranks: ragged blocks (3 + 2)
    W, (lo, hi) = frames_to_wigner_sharded(wl["psi"], wl["traced"], wl["xvec"], engine=eng, gather="nccl")
    Wl, _ = frames_to_wigner_sharded(wl["psi"], wl["traced"], wl["xvec"], engine=eng, gather=None)
    rho = workloads.random_rho(6, seed=2)
    x = np.linspace(-4, 4, 21)
    Ws, (rlo, rhi) = wigner_row_slabs(rho, x, engine=eng)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), lo=lo, hi=hi, rlo=rlo, rhi=rhi, local=Wl.numpy(),
             full=W.numpy() if W is not None else np.zeros(0), slab=Ws.numpy() if Ws is not None else np.zeros(0))
    dist.barrier()
    dist.destroy_process_group()


class _UnitRangeEngine:
    """Device-engine stand-in on CPU tensors: ``wigner_t`` with ``unit_range`` fills only the points of that share
    of an S x S mirror-symmetric grid's units (unit = pair (i <= j) of grid lines and its mirror images)."""

    device = "cpu"
    _grid_sharing = True

    def __init__(self, supported=True):
        import torch

        self.torch_device = torch.device("cpu")
        self.supported = supported
        self.calls = []

    def grid_is_mirror(self, xvec, yvec=None):
        x = np.asarray(xvec)
        return yvec is None and np.allclose(x, -x[::-1], rtol=0, atol=1e-15 * np.abs(x).max())

    def grid_units(self, S):
        U = (S + 1) // 2
        return U * (U + 1) // 2

    def wigner_t(self, rho, xvec, yvec=None, g=np.sqrt(2), method="clenshaw", out=None, unit_range=None):
        import torch

        from bosonic_qiskit_b200._lib import BQ_ERR_UNSUPPORTED, BQError
        from oracle import wigner_clenshaw

        self.calls.append(unit_range)
        x = xvec.numpy()
        y = x if yvec is None else yvec.numpy()
        W = wigner_clenshaw(rho.numpy(), x, y, g)
        if unit_range is None:
            return torch.from_numpy(W)
        if not self.supported:
            raise BQError(BQ_ERR_UNSUPPORTED, "no symmetric-grid kernel")
        S = len(x)
        U = (S + 1) // 2
        units = [(i, j) for i in range(U) for j in range(i, U)]
        o = out[0].numpy()
        for i, j in units[unit_range[0]: unit_range[0] + unit_range[1]]:
            for a, b in ((i, j), (j, i)):
                for aa in (a, S - 1 - a):
                    for bb in (b, S - 1 - b):
                        o[aa, bb] = W[aa, bb]
        return out


def _unit_worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist

    import workloads
    from bosonic_qiskit_b200.parallel import wigner_row_slabs

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rho = workloads.random_rho(6, seed=2)
    res = {}
    # mirror grid: shared out by unit range + one reduce; then an asymmetric grid (row slabs + gather) right after it,
    # so a rank that took a different branch (an unmatched collective) shows up as a wrong or hung second result
    for name, x, supported in (("mirror", np.linspace(-4, 4, 21), True), ("skew", np.linspace(-3, 4, 20), True),
                               ("refused", np.linspace(-4, 4, 12), False), ("mirror2", np.linspace(-2, 2, 16), True)):
        eng = _UnitRangeEngine(supported)
        W, _ = wigner_row_slabs(rho, x, engine=eng)
        res[name] = W.numpy() if W is not None else np.zeros(0)
        res[name + "_calls"] = np.array([-1 if c is None else c[1] for c in eng.calls])
    np.savez(os.path.join(out_dir, f"unit_rank{rank}.npz"), **res)
    dist.barrier()
    dist.destroy_process_group()


def test_shared_units_then_row_slabs_gloo(tmp_path):
    """Every rank must take the same branch of ``wigner_row_slabs`` (mirror grid -> unit ranges + reduce; otherwise
    row slabs + gather), including the ranks that receive no result."""
    import torch.multiprocessing as mp

    import workloads
    from oracle import wigner_clenshaw

    world, port = 2, 31500 + os.getpid() % 2000
    mp.spawn(_unit_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "unit_rank0.npz"), np.load(tmp_path / "unit_rank1.npz")
    rho = workloads.random_rho(6, seed=2)
    for name, x in (("mirror", np.linspace(-4, 4, 21)), ("skew", np.linspace(-3, 4, 20)),
                    ("refused", np.linspace(-4, 4, 12)), ("mirror2", np.linspace(-2, 2, 16))):
        assert r1[name].size == 0
        assert np.abs(r0[name] - wigner_clenshaw(rho, x)).max() < 1e-13, name
    for r in (r0, r1):
        assert len(r["mirror_calls"]) == 1 and r["mirror_calls"][0] >= 0  # one unit-range launch, no slab launch
        assert list(r["skew_calls"]) == [-1]  # one slab launch
        assert len(r["refused_calls"]) == 2 and r["refused_calls"][1] == -1  # refused unit range, then the slab
    assert int(r0["mirror_calls"][0]) + int(r1["mirror_calls"][0]) == 11 * 12 // 2


def test_sharded_frames_and_row_slabs_gloo(tmp_path):
    import torch.multiprocessing as mp

    import workloads
    from oracle import partial_trace_sv, wigner_clenshaw

    world, port = 2, 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    wl = workloads.config3(frames=5, steps=24)
    ref = np.stack([wigner_clenshaw(partial_trace_sv(p, wl["traced"]), wl["xvec"]) for p in wl["psi"]])
    assert (int(r0["lo"]), int(r0["hi"]), int(r1["lo"]), int(r1["hi"])) == (0, 3, 3, 5)
    assert r0["full"].shape == (5, 24, 24) and r1["full"].size == 0  # gathered on rank 0 only
    assert np.abs(r0["full"] - ref).max() < 1e-13
    assert np.abs(r0["local"] - ref[:3]).max() < 1e-13 and np.abs(r1["local"] - ref[3:]).max() < 1e-13
    rho = workloads.random_rho(6, seed=2)
    x = np.linspace(-4, 4, 21)
    assert (int(r0["rlo"]), int(r0["rhi"]), int(r1["rlo"]), int(r1["rhi"])) == (0, 11, 11, 21)
    assert np.abs(r0["slab"] - wigner_clenshaw(rho, x)).max() < 1e-13
