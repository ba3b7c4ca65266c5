cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) 2>&1 | tail -12
gcc -O2 -pthread -Iinclude examples/ckzg_threads.c -Lrust-kzg_b200 -lb200kzg -Wl,-rpath,$PWD/rust-kzg_b200 -o /tmp/ckt || exit 1
S=rust-kzg_b200/data/trusted_setup.txt
OUT=gpurun_out/r2_threads_direct.jsonl
: > $OUT
for op in commit blob_proof; do for lanes in 1 2 4; do for t in 1 4 16 64; do
  echo -n "{\"lanes\": $lanes, \"run\": " >> $OUT
  B200_KZG_LANES=$lanes /tmp/ckt $S $op $t 150 4 | tr -d '\n' >> $OUT
  echo "}" >> $OUT
done; done; done
python - <<'PY'
import json
for l in open('gpurun_out/r2_threads_direct.jsonl'):
    d=json.loads(l); r=d['run']
    print(r['op'],'lanes',d['lanes'],'T',r['threads'],'per_s=%.0f'%r['per_s'],'batch=%.2f'%r['mean_batch'],'exec=%.0f'%r['mean_lane_exec_us'],'wait=%.0f'%r['mean_lane_wait_us'], 'bad', r['mismatches']+r['errors'])
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv -c 400 --log-file gpurun_out/r2_launches_blob1_direct.csv python scripts/ncu_target.py blob 1 2 > /dev/null 2>&1
python scripts/launch_table.py gpurun_out/r2_launches_blob1_direct.csv 2>/dev/null | tail -8
ncu --metrics gpu__time_duration.sum --clock-control none --csv -c 400 --log-file gpurun_out/r2_launches_blob16_direct.csv python scripts/ncu_target.py blob 16 2 > /dev/null 2>&1
python scripts/launch_table.py gpurun_out/r2_launches_blob16_direct.csv 2>/dev/null | tail -8
