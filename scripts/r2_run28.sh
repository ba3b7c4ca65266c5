cd $GRAFT_REPO_ROOT
timeout 600 python scripts/direct_bits_sweep.py 13 2>&1 | tail -2
for nb in 1 16 64; do B200_DIRECT_TRACE=1 python scripts/ncu_target.py blob $nb 4 2>&1 | grep "direct trace" | tail -1; done
