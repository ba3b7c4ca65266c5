cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_arith.py tests/test_gpu_verify.py tests/test_gpu_eip4844.py tests/test_gpu_das7594.py -m gpu -x -q 2>&1 | tail -3
python scripts/verify_timing.py 2>&1 | tail -6
gcc -O2 -pthread -Iinclude examples/ckzg_threads.c -Lrust-kzg_b200 -lb200kzg -Wl,-rpath,$PWD/rust-kzg_b200 -o /tmp/ckt || exit 1
/tmp/ckt rust-kzg_b200/data/trusted_setup.txt blob_proof 1 100 4 | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('blob_proof T1 per_s=%.0f'%r['per_s'],'bad',r['mismatches']+r['errors'])"
