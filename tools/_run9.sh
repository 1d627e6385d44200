cd $GRAFT_REPO_ROOT
timeout 300 python tools/phase_timeline.py arabidopsis 0 2>&1 | tail -9
timeout 300 python tools/phase_timeline.py arabidopsis 1 2>&1 | tail -9
