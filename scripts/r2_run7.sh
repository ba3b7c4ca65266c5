cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) 2>&1 | tail -12
timeout 600 python scripts/blob_window_sweep.py 12:-1 12:2 13:2 13:3 14:3 2>&1 | tail -8
cp gpurun_out/blob_window_sweep.json gpurun_out/r2_blob_window_sweep_randomized.json
(time timeout 900 python bench.py --steps 20 --warmup 5) > gpurun_out/r2_run7_bench.json 2> gpurun_out/r2_run7_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_run7_bench.json'))
print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'pageable', d['e2e_pageable']['ms_per_step'])
ex=d['extra']
print('adv msm', {k:(round(v['vs_uniform'],2), v['parity_ok']) for k,v in ex['adversarial_msm'].items() if isinstance(v,dict)})
print('adv blobs', {k:(round(v['vs_uniform'],2), v['parity_ok']) for k,v in ex['adversarial_blobs'].items() if isinstance(v,dict)})
print('commit', ex['blob_to_kzg_commitment']['ms_per_batch'], 'single', ex['blob_to_kzg_commitment']['single_blob_ms'], 'proof', ex['compute_kzg_proof']['ms_per_batch'])
print('threads', ex['threads'])
print('2p24', ex['msm_2p24'])
PY
