cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_das7594.py tests/test_gpu_eip4844.py -m gpu -x -q 2>&1 | tail -3
python scripts/fk20_timing.py 2>&1 | tail -6
B200_FFT_G1_FUSE=0 python scripts/fk20_timing.py 2>&1 | tail -6
python scripts/verify_timing.py 2>&1 | grep -E "compute_cells_and|recover_cells_and"
