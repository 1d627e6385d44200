cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 300 python tools/profile_scan.py arabidopsis 5 2>&1 | tail -3
timeout 300 python tools/phase_timeline.py arabidopsis 0 2>&1 | tail -9
CRP_WAVE_TILES=3000 timeout 300 python tools/profile_scan.py arabidopsis 3 2>&1 | tail -1
timeout 300 python tools/profile_scan.py sorghum 3 | tail -1
