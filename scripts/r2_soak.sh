# soak of the coalescers through the pthread driver: many more calls per thread than the tests make (run under gpurun)
cd $GRAFT_REPO_ROOT
gcc -O2 -pthread -Iinclude examples/ckzg_threads.c -Lrust-kzg_b200 -lb200kzg -Wl,-rpath,$PWD/rust-kzg_b200 -o /tmp/ckt || exit 1
S=rust-kzg_b200/data/trusted_setup.txt
run() { timeout 600 /tmp/ckt $S "$@" | python -c "import sys,json; r=json.loads(sys.stdin.read()); print(r['op'],'T',r['threads'],'calls',r['calls'],'per_s=%.0f'%r['per_s'],'batch=%.2f'%r['mean_batch'],'mismatches',r['mismatches'],'errors',r['errors'],'isolation',r['isolation_failures'])"; echo "rc=$?"; }
run commit 64 400 4
run mixed 48 200 4
run blob_proof 32 200 4
run verify 32 300 4
run verify 7 300 3
run cells 16 60 2
run cells 5 60 3
run cells 40 30 2
