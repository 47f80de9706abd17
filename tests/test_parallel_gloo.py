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
