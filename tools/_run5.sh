cd $GRAFT_REPO_ROOT
CRP_TRACE=1 timeout 300 python tools/profile_scan.py arabidopsis 3 2>&1 | tail -40
