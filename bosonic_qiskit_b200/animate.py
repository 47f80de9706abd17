"""Drop-in for the frame loops of ``bosonic_qiskit.animate``.

``animate_wigner`` keeps the reference signature (``src/bosonic_qiskit/animate.py:28-46``).  What
changes is where the per-frame analysis runs:

  * without ``qubit``/``cbit`` (``animate.py:211-247``): one Aer run saves a statevector per segment;
    the loop ``trace_out_qubits -> _wigner`` per frame (``wigner.py:92-96``) becomes one batched GPU
    call (``wigner.simulate_wigner_multiple_statevectors``);
  * with ``qubit``/``cbit`` and ``keep_state=True`` (``animate.py:314-354``): frame k is simulated from
    frame k-1's statevector, so the Aer runs stay sequential, but the Wigner grid is not fed back --
    the statevectors are collected and analysed in one batched call after the loop;
  * ``keep_state=False`` (``animate.py:178-206``): the reference fans frames out over a
    ``multiprocessing.Pool``; here the simulations run in-process and the analysis is batched (the
    reference passes a non-existent ``conditional=`` kwarg / a bool ``conditional_state`` in these
    branches -- latent bugs; the intended "0x0" conditional state is used instead).

Circuit discretisation, Aer simulation and matplotlib rendering are the reference's own code,
imported lazily from ``bosonic_qiskit`` -- they are outside the GPU hot path.
"""

from __future__ import annotations

from typing import Sequence

import numpy as np

from . import wigner as _w
from .util import _find_qubit_indices

# test seams (None => the reference's implementation)
_discretize_single_circuit_impl = None
_discretize_circuits_impl = None


def _ref_discretize():
    if _discretize_single_circuit_impl is not None or _discretize_circuits_impl is not None:
        return _discretize_circuits_impl, _discretize_single_circuit_impl
    try:
        from bosonic_qiskit.discretize import discretize_circuits, discretize_single_circuit
    except Exception as exc:  # pragma: no cover - needs qiskit
        raise RuntimeError("animate_wigner needs the reference package `bosonic_qiskit` to discretise circuits") from exc
    return discretize_circuits, discretize_single_circuit


def wigner_frames_without_measure(
    circuit,
    animation_segments: int = 10,
    discretize_epsilon: float | None = None,
    shots: int = 1,
    axes_min: int = -6,
    axes_max: int = 6,
    axes_steps: int = 200,
    noise_passes=None,
    sequential_subcircuit: bool = False,
    trace: bool = True,
):
    """``__discretize_wigner_without_measure`` (``animate.py:211-247``): returns ``(w_fock, xvec)``."""
    statevector_label = "segment_"
    _, discretize_single_circuit = _ref_discretize()
    discretized, num_statevectors = discretize_single_circuit(
        circuit=circuit,
        segments_per_gate=animation_segments,
        epsilon=discretize_epsilon,
        sequential_subcircuit=sequential_subcircuit,
        statevector_per_segment=True,
        statevector_label=statevector_label,
        noise_passes=noise_passes,
    )
    xvec = np.linspace(axes_min, axes_max, axes_steps)
    w_fock = _w.simulate_wigner_multiple_statevectors(
        circuit=discretized,
        xvec=xvec,
        shots=shots,
        statevector_label=statevector_label,
        num_statevectors=num_statevectors,
        noise_passes=noise_passes,
        trace=trace,
    )
    return w_fock, xvec


def wigner_frames_with_state(circuits: Sequence, qubit, cbit, xvec, shots: int, noise_passes=None,
                             trace: bool = False) -> list:
    """``__simulate_wigner_with_state`` (``animate.py:314-354``) with the analysis deferred and batched."""
    frames = []
    previous_state = None
    traced = None
    do_trace = trace or cbit is not None
    for circuit in circuits:
        sim_circuit = circuit
        if previous_state:
            sim_circuit = circuit.copy_empty_like()
            sim_circuit.initialize(previous_state)
            last_instructions = circuit.data[-3:] if (qubit and cbit) else circuit.data[-1:]
            for inst in last_instructions:
                sim_circuit.append(*inst)
        conditional_state = "0x0" if cbit else None
        states, _, _ = _w._simulate(
            sim_circuit,
            shots=shots,
            noise_passes=noise_passes,
            conditional_state_vector=True if conditional_state else False,
            return_fockcounts=False,
        )
        if not states:
            raise RuntimeError("No state vector returned from the simulation")
        previous_state = states[conditional_state] if conditional_state else states
        if traced is None:
            traced = _find_qubit_indices(circuit) if do_trace else []
        frames.append(np.asarray(previous_state.data))
    if not frames:
        return []
    W = _w.frames_to_wigner(frames, traced, xvec)
    return [W[k] for k in range(W.shape[0])]


def wigner_frames_with_measure(
    circuit,
    qubit=None,
    cbit=None,
    animation_segments: int = 10,
    shots: int = 1,
    axes_min: int = -6,
    axes_max: int = 6,
    axes_steps: int = 200,
    processes: int | None = None,
    keep_state: bool = True,
    noise_passes=None,
    sequential_subcircuit: bool = False,
    trace: bool = True,
):
    """``__discretize_wigner_with_measure`` (``animate.py:146-208``): returns ``(w_fock, xvec)``.

    ``processes`` is accepted for signature compatibility; the GPU replaces the process pool."""
    discretize_circuits, _ = _ref_discretize()
    circuits = discretize_circuits(circuit, animation_segments, keep_state, qubit, cbit, sequential_subcircuit)
    xvec = np.linspace(axes_min, axes_max, axes_steps)
    if keep_state:
        w_fock = wigner_frames_with_state(circuits, qubit, cbit, xvec, shots, noise_passes, trace)
    else:
        frames, traced = [], None
        do_trace = trace or cbit is not None
        conditional_state = "0x0" if cbit is not None else None
        for frame_circuit in circuits:
            states, _, _ = _w._simulate(
                frame_circuit,
                shots=shots,
                noise_passes=noise_passes,
                conditional_state_vector=True if conditional_state else False,
                return_fockcounts=False,
            )
            if not states:
                continue  # the reference drops frames that produced no result (animate.py:112,206)
            state = states[conditional_state] if conditional_state else states
            if traced is None:
                traced = _find_qubit_indices(frame_circuit) if do_trace else []
            frames.append(np.asarray(state.data))
        W = _w.frames_to_wigner(frames, traced, xvec) if frames else np.empty((0, axes_steps, axes_steps))
        w_fock = [W[k] for k in range(W.shape[0])]
    return w_fock, xvec


def frames_to_rgba(frames, levels: int = 100, zero_abs_max: float = 5.0) -> np.ndarray:
    """RGBA images of Wigner frames without matplotlib: per frame the symmetric colour levels of ``_animate``
    (``animate.py:278-286``: ``linspace(-abs_max, abs_max, 100)``) and the ``RdBu`` colour ``contourf`` gives each
    grid node, computed on the GPU (min/max reduction + one colouring kernel, ``bq_frame_rgba_dev``).  Opt-in fast
    path for movie writers (``imshow(rgba, origin="lower")`` / raw frames to ffmpeg); ``animate_wigner`` itself keeps
    the reference's ``contourf`` rendering.  ``frames``: (F, S, S) or a list of (S, S) -> (F, S, S, 4) uint8."""
    import torch

    from .engine import get_engine

    eng = get_engine()
    arr = np.ascontiguousarray(np.stack([np.asarray(f, dtype=np.float64) for f in frames]) if not isinstance(frames, np.ndarray)
                               else frames, dtype=np.float64)
    single = arr.ndim == 2
    W = torch.from_numpy(arr[None] if single else arr).to(torch.device("cuda", eng.device))
    rgba = eng.frame_rgba_t(W, levels=levels, zero_abs_max=zero_abs_max).cpu().numpy()
    return rgba[0] if single else rgba


def animate_wigner(
    circuit,
    qubit=None,
    cbit=None,
    animation_segments: int = 10,
    discretize_epsilon: float | None = None,
    shots: int = 1,
    file=None,
    axes_min: int = -6,
    axes_max: int = 6,
    axes_steps: int = 200,
    processes: int | None = None,
    keep_state: bool = True,
    noise_passes=None,
    sequential_subcircuit: bool = False,
    draw_grid: bool = False,
    trace: bool = True,
    bitrate: int = -1,
):
    """Animate the Wigner function at each discretised step of ``circuit`` (``animate.py:28-143``).

    Frames are produced on the GPU; the matplotlib ``FuncAnimation`` and the movie writers are the
    reference's (``animate.py:116-141,250-311``)."""
    if qubit or cbit:
        w_fock, xvec = wigner_frames_with_measure(
            circuit, qubit, cbit, animation_segments, shots, axes_min, axes_max, axes_steps, processes, keep_state,
            noise_passes, sequential_subcircuit, trace,
        )
    else:
        w_fock, xvec = wigner_frames_without_measure(
            circuit, animation_segments, discretize_epsilon, shots, axes_min, axes_max, axes_steps, noise_passes,
            sequential_subcircuit, trace,
        )
    w_fock = [i for i in w_fock if i is not None]

    try:  # rendering: reference code + matplotlib, host side
        from functools import partial

        import matplotlib.animation
        import matplotlib.pyplot as plt
        from bosonic_qiskit.animate import _animate, _animate_init, save_animation
    except Exception as exc:  # pragma: no cover - needs matplotlib + the reference package
        raise RuntimeError("rendering the animation needs matplotlib and the reference package `bosonic_qiskit`") from exc

    fig, ax = plt.subplots(constrained_layout=True)
    animate_fn = partial(_animate, fig=fig, ax=ax, xvec=xvec, wigner_results=w_fock, file=file, draw_grid=draw_grid)
    anim = matplotlib.animation.FuncAnimation(
        fig=fig, init_func=_animate_init, func=animate_fn, frames=len(w_fock), interval=200, repeat=True
    )
    if file:
        save_animation(anim, file, bitrate)
    return anim


__all__ = [
    "animate_wigner",
    "wigner_frames_without_measure",
    "wigner_frames_with_measure",
    "wigner_frames_with_state",
    "frames_to_rgba",
]
