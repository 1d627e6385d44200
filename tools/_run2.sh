cd $GRAFT_REPO_ROOT
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v6.log 2>&1; echo "bench rc=$?"; tail -2 gpurun_out/bench_v6.log
timeout 300 python tools/profile_scan.py arabidopsis 3 > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_scan_score -s 1 -c 1 -f -o gpurun_out/prof_scan_v6 python tools/profile_scan.py arabidopsis 3 > gpurun_out/ncu_v6.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_v6.log
