"""Drop-in for the Wigner entry points of ``bosonic_qiskit.wigner``.

Same names, keyword arguments, defaults and return layout (``W[iy, ix]`` float64) as the
reference (``src/bosonic_qiskit/wigner.py:26-190``).  The delegation
``qutip.wigner(psi=Qobj(rho), xvec, yvec, g, method)`` at ``wigner.py:190`` is replaced by the
sm_100a Clenshaw kernel (K3, ``csrc/wigner_kernels.cu``); the per-frame loop at
``wigner.py:92-96`` becomes one fused, batched launch.  ``util.simulate`` (Aer) is untouched and is
imported from the reference package when a circuit has to be simulated.

Quirks of the reference that are kept on purpose:
  * ``wigner()`` on a raw 1-D array forms ``np.outer(state, state)`` WITHOUT conjugation
    (``wigner.py:130-131``);
  * ``_wigner`` evaluates ``if not yvec`` (``wigner.py:187``), so ``yvec`` is effectively ``xvec``;
  * several kept qumodes are fed to the Wigner routine as ONE mode of dimension ``cutoff**M``.
"""

from __future__ import annotations

from typing import Literal, Sequence

import numpy as np

from ._compat import DensityMatrix, Statevector, is_density_matrix, is_statevector
from .engine import get_engine, method_code
from .util import _find_qubit_indices, trace_out_qubits

# Available method names in Qutip (wigner.py:22-23)
WignerMethods = Literal["clenshaw", "iterative", "laguerre", "fft"]
WignerResult = tuple[np.ndarray, ...]

_SQRT2 = np.sqrt(2)

# test seam: replaced by tests that have no qiskit-aer; None => the reference's own simulate()
_simulate_impl = None


def _simulate(*args, **kwargs):
    if _simulate_impl is not None:
        return _simulate_impl(*args, **kwargs)
    try:
        from bosonic_qiskit.util import simulate  # the reference package, unchanged
    except Exception as exc:  # pragma: no cover - needs qiskit + qiskit-aer
        raise RuntimeError(
            "simulating a circuit needs the reference package `bosonic_qiskit` (qiskit + qiskit-aer); "
            "only the trace + Wigner stages run on the GPU"
        ) from exc
    return simulate(*args, **kwargs)


def _check_fft(method):
    """``method="fft"`` is served by its own entry point (``bq_wigner_fft_*``): like ``qutip.wigner`` it returns the
    tuple ``(W, yvec)`` on the p axis it derives from the x grid, wherever the reference would return ``W``."""
    return method_code(method) == 3


def simulate_wigner(
    circuit,
    xvec: np.ndarray,
    shots: int,
    noise_passes=None,
    conditional_state: str | None = None,
    trace: bool = False,
    method: WignerMethods = "clenshaw",
    g: float = _SQRT2,
):
    """Simulate the circuit, optionally trace out the qubits, and compute the Wigner function
    (``wigner.py:26-71``).  Returns ``(W, statevector)``."""
    states, _, _ = _simulate(
        circuit,
        shots=shots,
        noise_passes=noise_passes,
        conditional_state_vector=True if conditional_state else False,
        return_fockcounts=False,
    )
    if not states:
        raise RuntimeError("No state vector returned from the simulation")

    state = states[conditional_state] if conditional_state else states
    traced = _find_qubit_indices(circuit) if trace else []
    xvec = np.asarray(xvec, dtype=np.float64)
    W = get_engine().state_to_wigner(np.asarray(state.data), traced, xvec, None, g=g, method=method)
    return W, state


def simulate_wigner_multiple_statevectors(
    circuit,
    xvec,
    shots: int,
    statevector_label: str,
    num_statevectors: int,
    noise_passes=None,
    trace: bool = False,
    g: float = _SQRT2,
    method: WignerMethods = "clenshaw",
) -> list:
    """One simulation, then trace + Wigner on every saved statevector ``{label}{k}``
    (``wigner.py:74-98``).  The frame loop is a single batched GPU call."""
    state, result, _ = _simulate(circuit, shots=shots, noise_passes=noise_passes)
    if not result.results:
        raise RuntimeError("No statevector was returned from simulation.")

    data = result.data()
    frames = [np.asarray(getattr(data[f"{statevector_label}{num}"], "data", data[f"{statevector_label}{num}"]))
              for num in range(num_statevectors)]
    if not frames:
        return []
    traced = _find_qubit_indices(circuit) if trace else []
    W = frames_to_wigner(frames, traced, xvec, g=g, method=method)
    if _check_fft(method):  # one (W, yvec) tuple per frame, what the reference's loop would collect from qutip
        W, yvec = W
        return [(W[k], yvec) for k in range(W.shape[0])]
    return [W[k] for k in range(W.shape[0])]


# A frame loop goes to every visible GPU (one process, ``MultiEngine``) once it has at least this many frames per GPU
# -- below that one GPU finishes before the other devices' worker threads have copied their inputs in.
MULTI_GPU_MIN_FRAMES_PER_DEVICE = 4
# ... and ONE grid is shared out over the GPUs from this much Clenshaw work (flops) on: cutoff 1024 on 4096 x 4096 is
# 9e13 (0.4 s on one GPU), cutoff 64 on 1024 x 1024 is 2e10 (0.15 ms: not worth a second device).
MULTI_GPU_MIN_GRID_FLOPS = 5e12


def _frame_devices(devices, n_frames: int):
    """Devices a frame loop runs on: ``devices`` as given, or (None) every visible GPU when there are enough frames."""
    from .engine import device_count

    if devices is None:
        n = device_count()
        if n > 1 and n_frames >= MULTI_GPU_MIN_FRAMES_PER_DEVICE * n:
            return list(range(n))
        return None
    devices = [int(d) for d in devices]
    return devices if len(devices) > 1 else None


def frames_to_wigner(frames, traced: Sequence[int], xvec, g: float = _SQRT2, method: WignerMethods = "clenshaw",
                     out=None, devices=None) -> np.ndarray:
    """Batched body of the reference's frame loops: (F, 2**n) statevectors -> (F, S, S) Wigner grids.

    Replaces ``for num in range(num_statevectors): trace_out_qubits(...); _wigner(...)``
    (``wigner.py:92-96``, ``animate.py:326-352``) and the ``multiprocessing.Pool`` fan-out of
    ``animate.py:191-206`` with one call: ``bq_state_to_wigner_host`` on one GPU, or -- ``devices=None`` with several
    GPUs visible and enough frames, or an explicit list of devices -- ``bq_state_to_wigner_sharded_host``, which splits
    the frames over the GPUs from this one process and lets each GPU write its grids into its slice of the result.
    """
    psi = np.ascontiguousarray(np.stack([np.asarray(f, dtype=np.complex128).reshape(-1) for f in frames]))
    xvec = np.asarray(xvec, dtype=np.float64)
    devs = None if _check_fft(method) else _frame_devices(devices, psi.shape[0])
    if devs is not None:
        from .engine import get_multi_engine

        return get_multi_engine(devs).state_to_wigner(psi, list(traced), xvec, None, g=g, method=method, out=out)
    eng = get_engine(None if not devices else int(list(devices)[0]))
    return eng.state_to_wigner(psi, list(traced), xvec, None, g=g, method=method, out=out)


def wigner(
    state,
    axes_min: int = -6,
    axes_max: int = 6,
    axes_steps: int = 200,
    g: float = _SQRT2,
    method: WignerMethods = "clenshaw",
):
    """Wigner function of a state on ``linspace(axes_min, axes_max, axes_steps)**2`` (``wigner.py:101-136``)."""
    if is_statevector(state):
        state = DensityMatrix(state)

    if not is_density_matrix(state):
        state = np.atleast_1d(state)

        # If we have a statevector, turn it into density matrix
        if state.ndim == 1:
            state = np.outer(state, state)  # NB: no conjugate, as in the reference (wigner.py:130-131)

        state = DensityMatrix(state)

    xvec = np.linspace(axes_min, axes_max, axes_steps)
    return _wigner(state, xvec, g=g, method=method)


def wigner_mle(
    states: Sequence,
    axes_min: int = -6,
    axes_max: int = 6,
    axes_steps: int = 200,
    g: float = _SQRT2,
    method: WignerMethods = "clenshaw",
):
    """Maximum-likelihood (per-amplitude normal fit = mean) state over shots, then its Wigner function
    (``wigner.py:139-175``).  ``scipy.stats.norm.fit`` returns the sample mean as its location, which
    is what is computed here directly."""
    states_data = np.stack([np.asarray(getattr(s, "data", s)) for s in states])
    mle_state = states_data.mean(axis=0)
    mle_normalized = mle_state / np.linalg.norm(mle_state)
    return wigner(mle_normalized, axes_min, axes_max, axes_steps, g=g, method=method)


def _wigner(state, xvec, yvec=None, g: float = _SQRT2, method: WignerMethods = "clenshaw"):
    """``qutip.wigner(Qobj(rho), xvec, yvec, g, method)`` on the GPU (``wigner.py:178-190``)."""
    rho = np.asarray(state.data)

    if not yvec:  # same truthiness test as the reference (wigner.py:187)
        yvec = xvec

    if rho.ndim == 1:  # qutip turns kets into density matrices
        rho = np.outer(rho, rho.conj())
    if not _check_fft(method) and rho.ndim == 2:
        from .engine import device_count

        N, S = rho.shape[0], len(xvec)
        if 5.0 * N * N * S * len(yvec) >= MULTI_GPU_MIN_GRID_FLOPS and device_count() > 1:
            from .engine import get_multi_engine

            # one huge grid: shared out over every GPU of the box (unit ranges / row slabs), assembled on the host
            return get_multi_engine().wigner(rho, np.asarray(xvec, dtype=np.float64), np.asarray(yvec, dtype=np.float64),
                                             g=g, method=method, shard="grid")
    return get_engine().wigner(rho, np.asarray(xvec, dtype=np.float64), np.asarray(yvec, dtype=np.float64), g=g,
                               method=method)


def plot_wigner(
    circuit,
    state_vector,
    trace: bool = True,
    file=None,
    axes_min: int = -6,
    axes_max: int = 6,
    axes_steps: int = 200,
    num_colors: int = 100,
    draw_grid: bool = False,
    dpi: int = 100,
    **wigner_kwargs,
):
    """Wigner function of ``state_vector`` (traced to the qumodes by default) handed to ``plot``
    (``wigner.py:193-244``).  Rendering stays on the host (matplotlib); the trace + Wigner stages run on the GPU."""
    state = trace_out_qubits(circuit, state_vector) if trace else DensityMatrix(state_vector)
    w_fock = wigner(state=state, axes_min=axes_min, axes_max=axes_max, axes_steps=axes_steps, **wigner_kwargs)
    plot(data=w_fock, axes_min=axes_min, axes_max=axes_max, axes_steps=axes_steps, file=file, num_colors=num_colors,
         draw_grid=draw_grid, dpi=dpi)


def _pyplot():
    try:
        import matplotlib.pyplot as plt
        import matplotlib.ticker as tick
    except Exception as exc:  # pragma: no cover - matplotlib is a rendering dependency only
        raise RuntimeError("plotting needs matplotlib (rendering is host-side; only trace + Wigner run on the GPU)") from exc
    return plt, tick


def plot(data, axes_min: int = -6, axes_max: int = 6, axes_steps: int = 200, file=None, num_colors: int = 100,
         draw_grid: bool = False, dpi: int = 100):
    """Contour plot of a Wigner grid (``wigner.py:247-287``): symmetric colour levels
    ``linspace(-A, A, num_colors)``, ``A = max(amax, |amin|)`` (1 for an all-zero grid), ``contourf(..., cmap="RdBu")``,
    colour bar labelled W(x,p).  Host-side matplotlib, as in the reference."""
    plt, tick = _pyplot()
    axis = np.linspace(axes_min, axes_max, axes_steps)
    hi, lo = np.amax(data), np.amin(data)
    if hi == 0 and lo == 0:
        hi, lo = 1, -1
    bound = max(hi, abs(lo))
    levels = np.linspace(-bound, bound, num_colors)
    fig, ax = plt.subplots(constrained_layout=True)
    filled = ax.contourf(axis, axis, data, levels, cmap="RdBu")
    ax.set_xlabel(r"$x$")
    ax.set_ylabel(r"$p$")
    ax.set_aspect("equal", "box")
    if draw_grid:
        ax.grid()
    bar = fig.colorbar(filled, ax=ax, format=tick.FormatStrFormatter("%.2f"))
    bar.set_label(r"$W(x,p)$", rotation=270, labelpad=25)
    if file:
        plt.savefig(file, dpi=dpi)
    else:
        plt.show()


def snapshot_wigners(circuit, result, trace: bool = True, axes_min: int = -6, axes_max: int = 6, axes_steps: int = 200,
                     **wigner_kwargs):
    """The Wigner grids of every ``cv_snapshot_{k}`` statevector of a result, as ONE batched trace -> Wigner call
    (the loop body of ``plot_wigner_snapshot``, ``wigner.py:378-410``, is one more frame loop).  Returns
    ``[(label, W), ...]``."""
    data = result.data()
    labels = [f"cv_snapshot_{k}" for k in range(circuit.cv_snapshot_id)]
    if not labels:
        return []
    frames = [np.asarray(getattr(data[label], "data", data[label])) for label in labels]
    method = wigner_kwargs.pop("method", "clenshaw")
    g = wigner_kwargs.pop("g", _SQRT2)
    if wigner_kwargs:
        raise TypeError(f"wigner() got an unexpected keyword argument {next(iter(wigner_kwargs))!r}")
    xvec = np.linspace(axes_min, axes_max, axes_steps)
    traced = _find_qubit_indices(circuit) if trace else []
    W = frames_to_wigner(frames, traced, xvec, g=g, method=method)
    if _check_fft(method):
        W, yvec = W
        return [(label, (W[k], yvec)) for k, label in enumerate(labels)]
    return [(label, W[k]) for k, label in enumerate(labels)]


def plot_wigner_snapshot(circuit, result, folder=None, trace: bool = True, axes_min: int = -6, axes_max: int = 6,
                         axes_steps: int = 200, num_colors: int = 100, **wigner_kwargs):
    """One PNG per ``cv_snapshot_{k}`` of ``result`` (``wigner.py:378-410``): the grids come from one batched GPU
    call (``snapshot_wigners``), each is then drawn by ``plot``."""
    from pathlib import Path

    for label, w_fock in snapshot_wigners(circuit, result, trace, axes_min, axes_max, axes_steps, **wigner_kwargs):
        file = Path(folder) / f"{label}.png" if folder else f"{label}.png"
        if isinstance(w_fock, tuple):  # method="fft": (W, yvec)
            w_fock = w_fock[0]
        plot(data=w_fock, axes_min=axes_min, axes_max=axes_max, axes_steps=axes_steps, file=file, num_colors=num_colors)


def wigner_projections(circuit, x, y_z, y_x, xvec=None, **wigner_kwargs):
    """The four Wigner grids of ``plot_wigner_projection`` (``wigner.py:290-375``): projections of the
    qumode state onto qubit 0 / 1 (Pauli Z) and + / - (Pauli X), computed the way the reference does it.

    ``x``: statevector of the circuit; ``y_z`` / ``y_x``: statevectors after an extra Z / X on the qubit.
    The reference forms ELEMENT-WISE products ``x.data * conj(x.data)`` etc. (1-D arrays of length 2**n,
    ``wigner.py:310,317-320``), hands each to ``trace_out_qubits`` as if it were a state, and combines the
    four reduced matrices with signs ``(+,+,+,+)/4`` and ``(+,-,-,+)/4`` (``:327-328``).  That is mirrored
    literally: the eight traces are ONE batched partial-trace launch, the four grids ONE batched Wigner launch.
    Returns ``(W_zero, W_one, W_plus, W_minus)``.
    """
    xd = np.asarray(getattr(x, "data", x), dtype=np.complex128).reshape(-1)
    traced = _find_qubit_indices(circuit)
    eng = get_engine()
    stack = []
    for y in (y_z, y_x):
        yd = np.asarray(getattr(y, "data", y), dtype=np.complex128).reshape(-1)
        xT, yT = xd.conj(), yd.conj()
        stack += [xd * xT, xd * yT, yd * xT, yd * yT]
    rho = eng.partial_trace_sv(np.stack(stack), traced)  # (8, D, D)
    proj = []
    for base in (0, 4):
        t_xx, t_xy, t_yx, t_yy = rho[base:base + 4]
        proj.append((t_xx + t_xy + t_yx + t_yy) / 4)
        proj.append((t_xx - t_xy - t_yx + t_yy) / 4)
    if xvec is None:
        xvec = np.linspace(-6, 6, 200)
    method = wigner_kwargs.pop("method", "clenshaw")
    g = wigner_kwargs.pop("g", _SQRT2)
    if wigner_kwargs:
        raise TypeError(f"_wigner() got an unexpected keyword argument {next(iter(wigner_kwargs))!r}")
    W = eng.wigner(np.stack(proj), np.asarray(xvec, dtype=np.float64), None, g=g, method=method)
    if _check_fft(method):
        W, yvec = W
        return (W[0], yvec), (W[1], yvec), (W[2], yvec), (W[3], yvec)
    return W[0], W[1], W[2], W[3]


def _add_contourf(ax, fig, title, x, y, z, draw_grid: bool = False):
    """One panel of ``plot_wigner_projection`` (``wigner.py:413-444``): symmetric colour levels from the panel's own
    extremes (at least 1e-4 wide), integer axis ticks."""
    bound = max(np.amax(z), abs(np.amin(z)), 0.0001)
    filled = ax.contourf(x, y, z, np.linspace(-bound, bound, 100), cmap="RdBu")
    ax.set_xlabel("x")
    ax.set_xticks(sorted({int(v) for v in x}))
    ax.set_ylabel("p")
    ax.set_yticks(sorted({int(v) for v in y}))
    if draw_grid:
        ax.grid()
    ax.set_title(title)
    fig.colorbar(filled, ax=ax)


def plot_wigner_projection(circuit, qubit, file=None, draw_grid: bool = False, **wigner_kwargs):
    """Plot the projections onto 0, 1, +, - of a one-qubit CVCircuit (``wigner.py:290-375``).

    The three simulations and the matplotlib figure are the reference's; the eight partial traces and the four
    Wigner grids run on the GPU (``wigner_projections``)."""
    x, _, _ = _simulate(circuit)
    circuit.z(qubit)
    y_z, _, _ = _simulate(circuit)
    circuit.data.pop()
    circuit.x(qubit)
    y_x, _, _ = _simulate(circuit)
    circuit.data.pop()
    xvec = np.linspace(-6, 6, 200)
    w0, w1, wp, wm = wigner_projections(circuit, x, y_z, y_x, xvec, **wigner_kwargs)
    plt, _ = _pyplot()
    fig, ((ax0, ax1), (ax2, ax3)) = plt.subplots(2, 2, figsize=(12.8, 12.8))
    _add_contourf(ax0, fig, "Projection onto zero", xvec, xvec, w0, draw_grid)
    _add_contourf(ax1, fig, "Projection onto one", xvec, xvec, w1, draw_grid)
    _add_contourf(ax2, fig, "Projection onto plus", xvec, xvec, wp, draw_grid)
    _add_contourf(ax3, fig, "Projection onto minus", xvec, xvec, wm, draw_grid)
    if file:
        plt.savefig(file)
    else:
        plt.show()


__all__ = [
    "wigner_projections",
    "plot_wigner_projection",
    "WignerMethods",
    "WignerResult",
    "simulate_wigner",
    "simulate_wigner_multiple_statevectors",
    "frames_to_wigner",
    "wigner",
    "wigner_mle",
    "_wigner",
    "plot",
    "plot_wigner",
    "plot_wigner_snapshot",
    "snapshot_wigners",
]
