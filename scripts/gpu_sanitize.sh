#!/bin/bash
# compute-sanitizer over a subset of the parity tests: memcheck (out-of-bounds / misaligned accesses) and racecheck
# (shared-memory hazards of vscan / events / emit). Logs under gpurun_out/.
mkdir -p gpurun_out
sel='golden_all_pixels or (small_random and (0 or 7 or 13 or 21 or 34)) or sparse_ids or unmerged_neighbours_match_merged or multi_call or vscan_row_segments or emit_column_segments or (large_radius and (130-96-128-12 or 257-64))'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 86 python -m pytest tests/test_parity_gpu.py -q -x -k "$sel" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|misaligned" gpurun_out/memcheck.log | head -8
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 87 python -m pytest tests/test_parity_gpu.py -q -x -k "golden_all_pixels or (small_random and (7 or 21)) or sparse_ids" > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|passed|failed|hazard|Race" gpurun_out/racecheck.log | head -12
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 86 python __graft_entry__.py --smoke > gpurun_out/memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"
grep -E "ERROR SUMMARY|smoke ok" gpurun_out/memcheck_smoke.log | head -4
