# thread / lane sweep of the coalesced single-blob entry points (examples/ckzg_threads.c); results -> gpurun_out/r2_threads_sweep.jsonl
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
gcc -O2 -pthread -Iinclude examples/ckzg_threads.c -Lrust-kzg_b200 -lb200kzg -Wl,-rpath,$PWD/rust-kzg_b200 -o /tmp/ckt || exit 1
S=rust-kzg_b200/data/trusted_setup.txt
OUT=gpurun_out/r2_threads_sweep.jsonl
: > $OUT
for lanes in 1 2 4; do
  for t in 1 4 16 64; do
    echo -n "{\"lanes\": $lanes, \"run\": " >> $OUT
    B200_KZG_LANES=$lanes /tmp/ckt $S commit $t 150 4 | tr -d '\n' >> $OUT
    echo "}" >> $OUT
  done
done
for cap in 1 4; do
  echo -n "{\"lanes\": 4, \"coalesce_cap\": $cap, \"run\": " >> $OUT
  B200_KZG_COALESCE=$cap /tmp/ckt $S commit 16 150 4 | tr -d '\n' >> $OUT
  echo "}" >> $OUT
done
echo -n "{\"lanes\": 4, \"run\": " >> $OUT; /tmp/ckt $S blob_proof 16 150 4 | tr -d '\n' >> $OUT; echo "}" >> $OUT
echo -n "{\"lanes\": 4, \"run\": " >> $OUT; /tmp/ckt $S blob_proof 64 100 4 | tr -d '\n' >> $OUT; echo "}" >> $OUT
cat $OUT
