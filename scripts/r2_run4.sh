cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_arith.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python scripts/affine_check.py 20 > gpurun_out/r2_affine_check2.json 2> gpurun_out/r2_affine_check2.err
echo "rc=$?" >> gpurun_out/r2_affine_check2.err
tail -3 gpurun_out/r2_affine_check2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_affine_check2.json'))
bad=[k for k,v in d['parity'].items() if not v['ok']]
print('parity cases', len(d['parity']), 'bad', bad)
for k,v in d['timing'].items(): print(k, {a:b for a,b in v.items() if a in ('ms','accumulate_ms','parity_ok','path','error')})
PY
B200_MSM_AFFINE_TREE=1 timeout 900 python scripts/affine_check.py 20 > gpurun_out/r2_affine_check2_tree.json 2>> gpurun_out/r2_affine_check2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_affine_check2_tree.json'))
bad=[k for k,v in d['parity'].items() if not v['ok']]
print('TREE parity cases', len(d['parity']), 'bad', bad)
for k,v in d['timing'].items(): print('TREE', k, {a:b for a,b in v.items() if a in ('ms','accumulate_ms','parity_ok','path','error')})
PY
