cd $GRAFT_REPO_ROOT
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3; echo "smoke rc=$?"
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 300 python tools/profile_scan.py arabidopsis 5 > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_scan_score -s 1 -c 1 -f -o gpurun_out/prof_scan_v8 python tools/profile_scan.py arabidopsis 3 > gpurun_out/ncu_v8.log 2>&1; echo "ncu rc=$?"; cat gpurun_out/plain.log
timeout 300 python tools/profile_scan.py arabidopsis 3 1 | tail -1
timeout 300 python tools/profile_scan.py sorghum 3 | tail -1
