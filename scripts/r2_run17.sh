cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eip4844.py tests/test_gpu_threads.py tests/test_c_consumer.py -m gpu -x -q 2>&1 | tail -3
gcc -O2 -pthread -Iinclude examples/ckzg_threads.c -Lrust-kzg_b200 -lb200kzg -Wl,-rpath,$PWD/rust-kzg_b200 -o /tmp/ckt || exit 1
S=rust-kzg_b200/data/trusted_setup.txt
for t in 1 4 16 64; do /tmp/ckt $S commit $t 200 4 | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('commit T',r['threads'],'per_s=%.0f'%r['per_s'],'batch=%.2f'%r['mean_batch'],'exec=%.0f'%r['mean_lane_exec_us'],'bad',r['mismatches']+r['errors'])"; done
ncu --metrics gpu__time_duration.sum --clock-control none --csv -c 400 --log-file gpurun_out/r2_l1.csv python scripts/ncu_target.py blob 1 2 > /dev/null 2>&1
python scripts/launch_table.py gpurun_out/r2_l1.csv k_blob_to_fr | tail -6
ncu --metrics gpu__time_duration.sum --clock-control none --csv -c 400 --log-file gpurun_out/r2_l4.csv python scripts/ncu_target.py blob 4 2 > /dev/null 2>&1
python scripts/launch_table.py gpurun_out/r2_l4.csv k_blob_to_fr | tail -6
