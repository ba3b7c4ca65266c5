cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_ntt.py -m gpu -x -q -k "cluster or matches_oracle or das" 2>&1 | tail -5
for cl in 16 8 0; do echo "cluster $cl: $(B200_NTT_CLUSTER=$cl python scripts/ntt_timing.py 2>&1 | tail -1)"; done
echo "auto: $(python scripts/ntt_timing.py 2>&1 | tail -1)"
