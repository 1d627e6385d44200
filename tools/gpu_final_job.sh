#!/bin/bash
# One gpurun call at the end of a change: the whole GPU suite, smoke(), the profile job (plain bench, phase
# timeline, e2e trace, ncu launch list, ncu full capture of the scan kernel) and the default bench line.
#   gpurun --timeout 1500 -- 'bash tools/gpu_final_job.sh r2h'
R=${1:-r2}
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/${R}_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
bash tools/gpu_profile_job.sh $R
echo "== default bench"; timeout 600 python bench.py > gpurun_out/${R}_bench_default.log 2> gpurun_out/${R}_bench_default.err; echo "rc=$?"; tail -1 gpurun_out/${R}_bench_default.log | cut -c1-1500
