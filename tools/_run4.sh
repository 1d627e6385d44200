cd $GRAFT_REPO_ROOT
echo "--- noscore"; timeout 300 python tools/profile_scan.py arabidopsis 4 1 2>&1 | tail -3
echo "--- waves 4120"; CRP_WAVE_TILES=4120 timeout 300 python tools/profile_scan.py arabidopsis 4 2>&1 | tail -2
echo "--- waves 2060"; CRP_WAVE_TILES=2060 timeout 300 python tools/profile_scan.py arabidopsis 4 2>&1 | tail -2
echo "--- sorghum"; timeout 300 python tools/profile_scan.py sorghum 3 2>&1 | tail -2
echo "--- sorghum 1 wave"; CRP_WAVE_TILES=18944 timeout 300 python tools/profile_scan.py sorghum 3 2>&1 | tail -2
echo "--- maize"; timeout 600 python tools/profile_scan.py maize 3 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v7.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_v7.log | cut -c1-600
