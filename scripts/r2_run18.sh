cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for nb in 1 2 4 8 16; do B200_DIRECT_TRACE=1 python scripts/ncu_target.py blob $nb 6 2>&1 | grep "direct trace" | tail -3; done
