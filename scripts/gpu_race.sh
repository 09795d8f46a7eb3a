#!/bin/bash
# racecheck with all reports printed; the producer -> consumer hand-over of emit_kernel goes through mbarriers, which
# racecheck does not model: those pairs (ring store vs load_counts) are filtered out, anything else is listed.
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 100000 python -m pytest tests/test_parity_gpu.py -q -x -k "golden_all_pixels or (small_random and (7 or 21 or 33)) or sparse_ids or (medium and 96-80) or (vscan_row_segments and (70-300-8-6 or 64-200)) or (emit_column_segments and (300-40 or 200-64))" > gpurun_out/racecheck_all.log 2>&1; echo "racecheck rc=$?"
grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/racecheck_all.log | tail -3
echo "distinct writer / reader sites:"
grep -E "Race reported between|and (Read|Write) access" gpurun_out/racecheck_all.log | sed -E 's/\+0x[0-9a-f]+//; s/\[[0-9]+ hazards\]//; s/\(shf::Geo[^)]*\)//; s/=========//; s/^[ .]*//' | sort | uniq -c | sort -rn | head -30
