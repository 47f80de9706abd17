"""Run under torchrun on >= 2 GPUs: checks the sharded frame loop and row-slab path on real devices.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29611 tests/multi_gpu_check.py

Every rank also computes the FULL result on its own GPU; rank 0 compares the gathered result with
it bit for bit (same kernel, same inputs => identical), and with the CPU oracle on one frame.
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def main():
    import torch
    import torch.distributed as dist

    import workloads
    from bosonic_qiskit_b200 import get_engine
    from bosonic_qiskit_b200.parallel import frames_to_wigner_sharded, wigner_row_slabs

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = get_engine(local)

    wl = workloads.config3(frames=2 * world + 1, steps=96)  # ragged blocks
    full_local = eng.state_to_wigner(wl["psi"], wl["traced"], wl["xvec"])
    ok = True
    for mode in ("nccl", "peer"):
        W, (lo, hi) = frames_to_wigner_sharded(wl["psi"], wl["traced"], wl["xvec"], engine=eng, gather=mode)
        if rank == 0:
            got = W.cpu().numpy()
            same = np.array_equal(got, full_local)
            print(f"[frames/{mode}] gathered {got.shape}, bit-identical to single-GPU result: {same}", flush=True)
            ok &= same
        else:
            assert W is None
        dist.barrier()

    rho = workloads.random_rho(96, seed=4, rank=3)
    x = np.linspace(-5, 5, 64)
    full_slab = eng.wigner(rho, x)
    Ws, (rlo, rhi) = wigner_row_slabs(rho, x, engine=eng)
    if rank == 0:
        # (a grid this small takes the point-by-point kernels on one GPU and the symmetric-grid kernels when it is
        # shared out by unit range: same function, different kernels -> compare to rounding, not bit for bit)
        err = float(np.abs(Ws.cpu().numpy() - full_slab).max() / np.abs(full_slab).max())
        print(f"[one grid, mirror-symmetric] assembled {tuple(Ws.shape)}, max|dW|/max|W| vs single GPU: {err:.2e}", flush=True)
        ok &= err < 1e-12
    xa = np.linspace(-4, 5, 64)  # not mirror-symmetric: row slabs of the point-by-point kernels + gather
    full_a = eng.wigner(rho, xa)
    Wl, (alo, ahi) = wigner_row_slabs(rho, xa, engine=eng, gather=None)  # this rank's slab, before any gather
    local_same = np.array_equal(Wl.cpu().numpy(), full_a[alo:ahi])
    print(f"[row slabs] rank {rank} rows [{alo}, {ahi}) bit-identical to its own full grid: {local_same}", flush=True)
    ok &= local_same
    Wa, _ = wigner_row_slabs(rho, xa, engine=eng)
    if rank == 0:
        got_a = Wa.cpu().numpy()
        same = np.array_equal(got_a, full_a)
        bad = np.argwhere(got_a != full_a)
        print(f"[row slabs] gathered {tuple(Wa.shape)}, bit-identical: {same}; max diff {np.abs(got_a - full_a).max():.2e}, "
              f"{len(bad)} points differ, first {bad[:3].tolist()}", flush=True)
        ok &= same
        from oracle import partial_trace_sv, wigner_clenshaw

        ref = wigner_clenshaw(partial_trace_sv(wl["psi"][1], wl["traced"]), wl["xvec"])
        err = np.abs(full_local[1] - ref).max() / np.abs(ref).max()
        print(f"[oracle] frame 1 max|dW|/max|W| = {err:.2e}", flush=True)
        ok &= err < 1e-10
    # one mirror-symmetric grid shared out by unit range + reduce (the symmetric-grid kernels)
    for Nc, Sc in ((32, 200), (100, 640)):
        rho2 = workloads.random_rho(Nc, seed=7, rank=3)
        x2 = np.linspace(-6, 6, Sc)
        full2 = np.array(eng.wigner(rho2, x2))
        Wu, _ = wigner_row_slabs(rho2, x2, engine=eng)
        if rank == 0:
            got = Wu.cpu().numpy()
            same = np.array_equal(got, full2)
            print(f"[shared units N={Nc} S={Sc}] reduced {got.shape}, bit-identical to single-GPU result: {same}, "
                  f"max diff {np.abs(got - full2).max():.2e}", flush=True)
            ok &= same
        dist.barrier()

    flag = torch.tensor([1 if ok else 0], device=torch.device("cuda", local))
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if int(flag.item()) else "FAIL", flush=True)
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
