#!/bin/bash
# compute-sanitizer passes over the smoke path (psi -> rho -> W at cutoff 16, then the ring kernel at cutoff 96): memcheck for
# out-of-bounds / misaligned accesses of the bulk-copy (cp.async.bulk) and mbarrier kernels, racecheck for shared-memory
# hazards between the producer thread's refills and the consuming warps, synccheck for the mbarrier / __syncwarp use.
#
#   tools/sanitize.sh [memcheck|racecheck|synccheck|initcheck ...]      (default: memcheck racecheck synccheck)
#
# Run on a GPU box you own:  on the shared B200 pool of this build compute-sanitizer is closed ("runs under it have left GPUs
# needing a reset"), so round 2 could not execute it there; the re-scheduled SASS is additionally compared bit for bit with
# nvcc's schedule when the library loads (bq_sass_self_check) and by tests/test_gpu_sass.py.
set -u
cd "$(dirname "$0")/.."
TOOLS=("$@")
[ ${#TOOLS[@]} -eq 0 ] && TOOLS=(memcheck racecheck synccheck)
rc=0
for mode in plain patched; do
  for tool in "${TOOLS[@]}"; do
    echo "== compute-sanitizer --tool $tool (BQ_B200_SASS=$mode)"
    BQ_B200_SASS=$mode BQ_B200_SASS_VERIFY=0 compute-sanitizer --tool "$tool" --error-exitcode 9 --print-limit 20 \
      python -c "import __graft_entry__ as g; g.smoke()" || rc=$?
  done
done
exit $rc
