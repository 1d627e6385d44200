#!/bin/bash
# usage (on the GPU box, via gpurun): bash tools/run_gpu.sh [tests] [bench]
cd ${GRAFT_REPO_ROOT:-.}
for what in "$@"; do
  case $what in
    tests) timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log;;
    bench) timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_latest.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_latest.log | cut -c1-900;;
    smoke) timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3;;
  esac
done
