cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) 2>&1 | tail -9
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err ) 2>&1 | tail -4
tail -3 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'points_per_s', 'gpu_launches')})
print('e2e', d['e2e'])
x = d.get('extra', {})
for k in ('blobs', 'threads', 'variable_base_e2e', 'adversarial_msm'):
    print(k, json.dumps(x.get(k))[:1500])
PY
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
