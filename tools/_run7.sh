cd $GRAFT_REPO_ROOT
timeout 300 python tools/profile_scan.py arabidopsis 3 1 > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_scan_score -s 1 -c 1 -f -o gpurun_out/prof_scan_v7_noscore python tools/profile_scan.py arabidopsis 3 1 > gpurun_out/ncu_v7n.log 2>&1; echo "ncu rc=$?"; cat gpurun_out/plain.log
