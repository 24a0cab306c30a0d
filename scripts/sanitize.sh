#!/bin/bash
# compute-sanitizer over the smoke scenes: memcheck (out-of-bounds / misaligned accesses, leaks),
# racecheck (shared-memory hazards: the scan kernels), synccheck (warp-synchronous votes / shuffles of the traversal,
# splat and grid kernels).  Run on a GPU box:  gpurun -- bash scripts/sanitize.sh   -> profiles/r2_sanitizer_*.log
# The smoke run is one Whitted render, one SPPM render and a ray-query batch on the "shadows" scene, each checked
# against the oracle, so a sanitizer-clean run is also a correct one.
set -u
cd "$(dirname "$0")/.."
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
SAN=${SANITIZER:-/usr/local/cuda/bin/compute-sanitizer}
rc=0
for tool in memcheck racecheck synccheck; do
    log="$OUT/r2_sanitizer_$tool.log"
    extra=""
    [ "$tool" = memcheck ] && extra="--leak-check full"
    timeout 900 "$SAN" --tool $tool $extra --error-exitcode 9 --log-file "$log" \
        python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/r2_sanitizer_$tool.out" 2>&1
    code=$?
    echo "== $tool: exit $code; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|LEAK SUMMARY' "$log" | tr '\n' ' ')"
    tail -1 "$OUT/r2_sanitizer_$tool.out"
    [ $code -ne 0 ] && rc=1
done
exit $rc
