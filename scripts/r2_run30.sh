cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err ) 2>&1 | tail -4
tail -3 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'points_per_s', 'gpu_launches')})
print('e2e', d['e2e'])
x = d.get('extra', {})
for k in x:
    if k not in ('fft_fr', 'das_fft_extension'):
        print(k, json.dumps(x[k])[:900])
PY
( time python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-600 ) 2>&1 | tail -5
