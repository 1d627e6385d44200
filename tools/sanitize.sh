#!/bin/bash
# compute-sanitizer over the scan kernel's small cases (on the GPU box, via gpurun):
# NOTE: on the GPU pool this repository was developed on, compute-sanitizer is closed (every tool answers with a
# refusal, profiles/r2_sanitizer_refused/); the script is kept for pools where it runs.  Stand-in: make checked.
#   bash tools/sanitize.sh            -> gpurun_out/sanitize_{memcheck,synccheck,racecheck,initcheck}.log
# Cases: smoke() (two segments, lane tables, both strands), the edge_fmt / multi3 golden CLIs, count
# ranges longer than one batch with 1-7 CTAs (ring refills, running prefix), static / ticketed tile
# dealing, tile-edge and dense-hit tiles.  Summaries are copied to profiles/ by hand.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
SEL='edge_fmt or test_long_count_ranges or test_static_and_ticketed or test_tile_boundaries or (test_cli_csv_is_byte_identical and multi3-)'
for tool in memcheck synccheck racecheck initcheck; do
  log=gpurun_out/sanitize_$tool.log
  { echo "### $CS --tool $tool  (smoke, then pytest -k \"$SEL\")";
    timeout 600 $CS --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15;
    timeout 1500 $CS --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" 2>&1 | tail -25; } > $log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tr '\n' ' ')"
done
