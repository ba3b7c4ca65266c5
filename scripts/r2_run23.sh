cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eip4844.py tests/test_gpu_threads.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python scripts/direct_bits_sweep.py 13 8 2>&1 | tail -3
for nb in 1 3 16 64; do B200_DIRECT_TRACE=1 python scripts/ncu_target.py blob $nb 4 2>&1 | grep "direct trace" | tail -1; done
gcc -O2 -pthread -Iinclude examples/ckzg_threads.c -Lrust-kzg_b200 -lb200kzg -Wl,-rpath,$PWD/rust-kzg_b200 -o /tmp/ckt || exit 1
S=rust-kzg_b200/data/trusted_setup.txt
for t in 1 16 64; do /tmp/ckt $S commit $t 200 4 | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('commit T',r['threads'],'per_s=%.0f'%r['per_s'],'batch=%.2f'%r['mean_batch'],'exec=%.0f'%r['mean_lane_exec_us'],'bad',r['mismatches']+r['errors'])"; done
