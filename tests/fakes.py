"""Minimal stand-ins for the reference's circuit objects (qiskit is not installed in the build image).

They reproduce exactly the index-layout contract the hot path depends on
(``src/bosonic_qiskit/qumoderegister.py:47-49,85-88`` and ``circuit.py:50-72,118-171``):
registers are laid out in constructor order, qumode k of a register with n qubits per mode owns n
consecutive qubit positions, and the Fock number is those bits read little-endian.
"""

from __future__ import annotations

from collections.abc import Sequence


class FakeQubit:
    _count = 0

    def __init__(self, label=""):
        FakeQubit._count += 1
        self.uid = FakeQubit._count
        self.label = label

    def __repr__(self):
        return f"Qubit({self.label}#{self.uid})"


class FakeQuantumRegister(Sequence):
    def __init__(self, size, name="q"):
        self.qubits = [FakeQubit(f"{name}{i}") for i in range(size)]

    def __getitem__(self, i):
        return self.qubits[i]

    def __len__(self):
        return len(self.qubits)


class FakeQumodeRegister(Sequence):
    def __init__(self, num_qumodes, num_qubits_per_qumode=2, name="qmr"):
        self.num_qumodes = num_qumodes
        self.num_qubits_per_qumode = num_qubits_per_qumode
        self.cutoff = 2 ** num_qubits_per_qumode
        self.qubits = [FakeQubit(f"{name}{i}") for i in range(num_qumodes * num_qubits_per_qumode)]

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self[i] for i in range(*k.indices(self.num_qumodes))]
        if k < 0:
            k += self.num_qumodes
        if not 0 <= k < self.num_qumodes:
            raise IndexError(k)
        n = self.num_qubits_per_qumode
        return self.qubits[k * n:(k + 1) * n]

    def __len__(self):
        return self.num_qumodes


class FakeCVCircuit:
    def __init__(self, *regs):
        self.qmregs = [r for r in regs if isinstance(r, FakeQumodeRegister)]
        self.qubits = []
        for r in regs:
            self.qubits.extend(r.qubits)

    @property
    def qumode_qubits(self):
        out = []
        for reg in self.qmregs:
            out.extend(reg.qubits)
        return out

    @property
    def qumode_qubit_indices(self):
        qm = set(self.qumode_qubits)
        return [i for i, q in enumerate(self.qubits) if q in qm]

    @property
    def qumode_qubits_indices_grouped(self):
        idx = {q: i for i, q in enumerate(self.qubits)}
        return [[idx[q] for q in qumode] for reg in self.qmregs for qumode in reg]

    def get_qubit_indices(self, qubits):
        flat = []
        for el in qubits:
            if isinstance(el, Sequence):
                flat += list(el)
            else:
                flat += [el]
        return [i for i, q in enumerate(self.qubits) if q in flat]
