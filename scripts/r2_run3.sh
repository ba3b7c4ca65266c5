cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python scripts/affine_check.py 20 > gpurun_out/r2_affine_check.json 2> gpurun_out/r2_affine_check.err
echo "rc=$?" >> gpurun_out/r2_affine_check.err
tail -3 gpurun_out/r2_affine_check.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_affine_check.json'))
bad=[k for k,v in d['parity'].items() if not v['ok']]
print('parity cases', len(d['parity']), 'bad', bad)
print({k:v['path'] for k,v in list(d['parity'].items())[:3]})
for k,v in d['timing'].items(): print(k, {a:b for a,b in v.items() if a in ('ms','accumulate_ms','parity_ok','path','error')})
PY
# lane / fill sweep of the coalescer
gcc -O2 -pthread -Iinclude examples/ckzg_threads.c -Lrust-kzg_b200 -lb200kzg -Wl,-rpath,$PWD/rust-kzg_b200 -o /tmp/ckt || exit 1
S=rust-kzg_b200/data/trusted_setup.txt
OUT=gpurun_out/r2_threads_fill.jsonl
: > $OUT
for fd in 1 2 4 8; do for lanes in 2 4; do for cap in 4 8 64; do
  echo -n "{\"fill_div\": $fd, \"lanes\": $lanes, \"cap\": $cap, \"run\": " >> $OUT
  B200_MSM_FILL_DIV=$fd B200_KZG_LANES=$lanes B200_KZG_COALESCE=$cap /tmp/ckt $S commit 16 120 4 | tr -d '\n' >> $OUT
  echo "}" >> $OUT
done; done; done
python - <<'PY'
import json
for l in open('gpurun_out/r2_threads_fill.jsonl'):
    d=json.loads(l); r=d['run']
    print('fd',d['fill_div'],'lanes',d['lanes'],'cap',d['cap'],'per_s=%.0f'%r['per_s'],'batch=%.2f'%r['mean_batch'],'exec=%.0f'%r['mean_lane_exec_us'],'wait=%.0f'%r['mean_lane_wait_us'])
PY
