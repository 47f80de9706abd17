#!/usr/bin/env python
"""Command-line entry of the SASS re-scheduler (the implementation is part of the package build)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bosonic_qiskit_b200._sass_resched import main  # noqa: E402

if __name__ == "__main__":
    main()
