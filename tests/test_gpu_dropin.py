"""The drop-in API cases (tests/dropin_cases.py) against the real CUDA engine."""

import inspect

import pytest

import dropin_cases as cases

pytestmark = pytest.mark.gpu
CASES = [getattr(cases, n) for n in sorted(dir(cases)) if n.startswith("case_")]


@pytest.mark.parametrize("case", CASES, ids=lambda f: f.__name__)
def test_dropin_on_gpu(case, monkeypatch, engine):
    before = engine.launches
    if "monkeypatch" in inspect.signature(case).parameters:
        case(monkeypatch)
    else:
        case()
    if case.__name__ != "case_get_qubit_indices":
        assert engine.launches > before, "the CUDA library did not run: silent fallback?"
