#!/bin/bash
# One GPU-box job (via gpurun): sanitizer pass, phase timeline, e2e trace, plain bench, ncu launch list + full capture.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
R=${1:-r2}
echo "== plain bench"; timeout 300 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/${R}_bench_plain.log 2> gpurun_out/${R}_bench_plain.err; echo "rc=$?"; tail -1 gpurun_out/${R}_bench_plain.log | cut -c1-600
echo "== phase timeline"; timeout 200 python tools/phase_timeline.py arabidopsis > gpurun_out/${R}_timeline.log 2>&1; cat gpurun_out/${R}_timeline.log | tail -12
echo "== e2e trace"; CRP_TRACE=1 timeout 200 python tools/e2e_trace.py arabidopsis > gpurun_out/${R}_e2e_trace.log 2>&1; grep -E "^rep" gpurun_out/${R}_e2e_trace.log
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/${R}_ncu_launches.log 2>&1; echo "rc=$?"
echo "== ncu full (scan)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_scan_score --launch-skip 1 --launch-count 1 -f -o gpurun_out/${R}_prof_scan python tools/profile_scan.py arabidopsis 3 > gpurun_out/${R}_ncu_full.log 2>&1; echo "rc=$?"; ls -la gpurun_out/${R}_prof_scan.ncu-rep
if [ "$2" = "sanitize" ]; then echo "== sanitizer"; bash tools/sanitize.sh; fi
