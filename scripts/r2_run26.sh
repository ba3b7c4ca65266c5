cd $GRAFT_REPO_ROOT
for t in 0 64 128 256 512 1024; do echo "tile $t: $(B200_NTT_TILE=$t python scripts/ntt_timing.py 2>&1 | tail -1)"; done
