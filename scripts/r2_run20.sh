cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for f in 0 1; do B200_FFT_G1_FUSE=$f python scripts/fk20_timing.py 2>&1 | tail -1; done
B200_FFT_G1_FUSE=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv -c 400 --log-file gpurun_out/r2_fk_fused.csv python scripts/ncu_target.py fk20 64 1 > /dev/null 2>&1
python scripts/launch_table.py gpurun_out/r2_fk_fused.csv k_blob_to_fr | tail -24
