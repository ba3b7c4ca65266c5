cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python scripts/fft_g1_batch_timing.py 2>&1 | tail -8
