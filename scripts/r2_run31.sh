cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_ntt.py -m gpu -x -q 2>&1 | tail -5
for cl in 1 0; do echo "cluster $cl: $(B200_NTT_CLUSTER=$cl python scripts/ntt_timing.py 2>&1 | tail -1)"; done
timeout 900 python -m pytest tests/test_gpu_eip4844.py tests/test_gpu_das7594.py -m gpu -x -q 2>&1 | tail -3
