"""Builds ``libbq_b200.so`` (the C-ABI library of include/bq_b200.h) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this also runs in the CPU-only build container.  Steps:

  1. ``wigner_kernels.cu`` -> ``csrc/_gen/wigner_kernels.cubin``                (nvcc -cubin, sm_100a)
  2. ``_sass_resched.py``   -> ``csrc/_gen/wigner_kernels.patched.cubin``       (DFMA operand-reuse chains)
  3. the patched image     -> ``csrc/_gen/wigner_cubin.inc``                   (C byte array)
  4. all ``.cu`` sources    -> ``libbq_b200.so`` with ``-DBQ_PATCHED_CUBIN``     (image embedded, loaded at run time)

If step 2 cannot patch anything the library is still built; it then runs nvcc's own schedule.
"""

from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
GEN = os.path.join(CSRC, "_gen")
LIB_PATH = os.path.join(_HERE, "libbq_b200.so")
SOURCES = ["bq_api.cu", "bq_multi.cu", "wigner_kernels.cu", "ptrace_kernels.cu", "wigner_fft_kernels.cu"]
HEADERS = ["bq_common.cuh", "bq_internal.h", os.path.join("..", "..", "include", "bq_b200.h"),
           os.path.join("..", "_sass_resched.py"), os.path.join("..", "_build.py")]

ARCH_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"]
NVCC_FLAGS = [*ARCH_FLAGS, "-Xcompiler", "-fPIC,-pthread", "-shared"]
MAX_REGS = 128  # the resident kernels' __launch_bounds__ (2 x 256 or 4 x 128 threads per SM) allow 128 registers
# the streaming kernels run one CTA per SM (their __launch_bounds__ say so): room for the re-scheduler's temporaries
MAX_REGS_FOR = [("wigner_stream_kernelILi4ELi256", 168), ("wigner_stream_kernelILi2ELi128", 168),
                ("wigner_resident_kernelILi8ELi128", 248), ("wigner_stream_kernelILi8ELi128", 248),
                ("wigner_diag_kernelILi8ELi256", 248), ("wigner_sym_resident_kernelILi3ELi128", 248),
                ("wigner_sym_diag_kernelILi3ELi256", 248),
                ("wigner_sym_resident_kernelILi2ELi128", 248), ("wigner_sym_resident_kernelILi2ELi192", 168),
                ("wigner_sym_diag_kernelILi2ELi256", 248), ("wigner_sym_diag_kernelILi2ELi384", 168),
                ("wigner_sym_diag_kernelILi1ELi256", 248),
                ("wigner_sym_resident_kernelILi4ELi256", 255), ("wigner_sym_resident_mixed_kernelILi4ELi3ELi256", 255), ("wigner_sym_diag_kernelILi4ELi256", 255),
                ("wigner_sym_diag_parked_kernelILi4ELi256", 248), ("wigner_sym_diag_parked_kernelILi3ELi384", 168)]


# Kernels whose straight-line FP64 blocks (the 8-accumulator tails of the symmetric-grid kernels) are re-scheduled as well
# (reschedule(block=True): 182 -> 153 issue cycles for the 74 FP64 instructions of the 2-unit tail): the 2-unit resident
# tiles.  Same box, same run (tools/sass_ab.sh, profiles/r2f_sass_ab.txt): 512 frames of cutoff 32 take 3.140 ms with nvcc's
# schedule, 2.967 ms with the loops re-scheduled, 2.833 ms with loops and tails.  Not the mixed 4 + 3 tile: only its 3-unit
# half is handled and c2 reads 0.1451 against 0.1439 ms with it; the tails of the other kernels the tool refuses (registers,
# or copies it cannot move).
BLOCKS_FOR = "wigner_sym_resident_kernelILi2ELi"


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libbq_b200.so (there is no CPU fallback)")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd, cwd):
    res = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("command failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return res.stdout + res.stderr


# The re-scheduler decodes the control words ptxas of THESE releases emits for sm_100a; under any other toolchain the
# build ships nvcc's own schedule (override: BQ_B200_FORCE_PATCH=1).  The library additionally compares every patched
# kernel with its nvcc-scheduled twin bit for bit when it is loaded (bq_sass_self_check).
PATCH_TOOLCHAINS = ("12.9.86",)


def nvcc_release() -> str:
    out = _run([_nvcc(), "--version"], CSRC)
    for tok in out.replace(",", " ").split():
        if tok.startswith("V") and tok[1:2].isdigit():
            return tok[1:]
    return "unknown"


def build_patched_image(verbose: bool = False) -> bool:
    """Steps 1-3.  Returns True when ``_gen/wigner_cubin.inc`` holds a re-scheduled image."""
    from . import _sass_resched

    os.makedirs(GEN, exist_ok=True)
    cubin = os.path.join(GEN, "wigner_kernels.cubin")
    patched = os.path.join(GEN, "wigner_kernels.patched.cubin")
    _run([_nvcc(), *ARCH_FLAGS, "-cubin", "-o", cubin, "wigner_kernels.cu"], CSRC)
    release = nvcc_release()
    if release not in PATCH_TOOLCHAINS and os.environ.get("BQ_B200_FORCE_PATCH") != "1":
        report = [f"nvcc {release} is not a toolchain the re-scheduler was validated with {PATCH_TOOLCHAINS}: "
                  "shipping nvcc's own schedule"]
        shutil.copyfile(cubin, patched)
    else:
        try:
            report = [f"nvcc {release}"] + _sass_resched.patch_file(cubin, patched, min_dfma=20, max_regs=MAX_REGS,
                                                                    max_regs_for=MAX_REGS_FOR, blocks_for=BLOCKS_FOR)
        except Exception as exc:  # the tool refuses rather than guesses; the refusal is in the report, and
            report = [f"re-scheduler failed: {exc!r}"]  # tests/test_sass_resched.py fails on it
            shutil.copyfile(cubin, patched)
    with open(os.path.join(GEN, "sass_resched_report.txt"), "w") as f:
        f.write("\n".join(report) + "\n")
    if verbose:
        print("\n".join(report))
    ok = any("FP64 issue model" in line for line in report)
    data = open(patched, "rb").read()
    with open(os.path.join(GEN, "wigner_cubin.inc"), "w") as f:
        f.write("// generated by bosonic_qiskit_b200/_build.py -- do not edit\n")
        f.write(f"static const unsigned long long kWignerCubinSize = {len(data)}ull;\n")
        f.write("alignas(16) static const unsigned char kWignerCubin[] = {\n")
        for i in range(0, len(data), 32):
            f.write(",".join(str(b) for b in data[i:i + 32]) + ",\n")
        f.write("};\n")
    return ok


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources into ``bosonic_qiskit_b200/libbq_b200.so``; returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    have_image = build_patched_image(verbose)
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", GEN, "-o", LIB_PATH, *SOURCES]
    if have_image:
        cmd.insert(1, "-DBQ_PATCHED_CUBIN")
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    out = _run(cmd, CSRC)
    if verbose:
        print(out)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
