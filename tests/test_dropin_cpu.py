"""Host logic of the drop-in wrappers, on CPU: the engine is replaced by the oracle-backed stand-in.

The same cases run against the real CUDA engine in tests/test_gpu_dropin.py."""

import pytest

import dropin_cases as cases
from oracle_engine import OracleEngine

CASES = [getattr(cases, n) for n in sorted(dir(cases)) if n.startswith("case_")]


@pytest.fixture(autouse=True)
def oracle_backed_engine(monkeypatch):
    eng = OracleEngine()
    import bosonic_qiskit_b200.engine as e
    import bosonic_qiskit_b200.util as u
    import bosonic_qiskit_b200.wigner as w

    for mod in (e, u, w):
        monkeypatch.setattr(mod, "get_engine", lambda device=None: eng)
    yield


@pytest.mark.parametrize("case", CASES, ids=lambda f: f.__name__)
def test_dropin_host_logic(case, monkeypatch):
    import inspect

    if "monkeypatch" in inspect.signature(case).parameters:
        case(monkeypatch)
    else:
        case()
