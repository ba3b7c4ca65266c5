# coalesced single verifications through the unmodified c-kzg symbol from 1..32 pthreads (run under gpurun)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_threads.py tests/test_gpu_verify.py tests/test_c_consumer.py -m gpu -x -q 2>&1 | tail -3
gcc -O2 -pthread -Iinclude examples/ckzg_threads.c -Lrust-kzg_b200 -lb200kzg -Wl,-rpath,$PWD/rust-kzg_b200 -o /tmp/ckt || exit 1
S=rust-kzg_b200/data/trusted_setup.txt
: > gpurun_out/r2_verify_threads.jsonl
for t in 1 4 8 16 32; do /tmp/ckt $S verify $t 30 4 | tee -a gpurun_out/r2_verify_threads.jsonl | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('verify T',r['threads'],'per_s=%.0f'%r['per_s'],'batch=%.2f'%r['mean_batch'],'bad',r['mismatches']+r['errors'])"; done
B200_KZG_VERIFY_COALESCE=1 /tmp/ckt $S verify 8 30 4 | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('uncoalesced T',r['threads'],'per_s=%.0f'%r['per_s'],'bad',r['mismatches']+r['errors'])"
