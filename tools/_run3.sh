cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/profile_scan.py arabidopsis 5 > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_scan_score -s 1 -c 1 -f -o gpurun_out/prof_scan_v7 python tools/profile_scan.py arabidopsis 3 > gpurun_out/ncu_v7.log 2>&1; echo "ncu rc=$?"; cat gpurun_out/plain.log
