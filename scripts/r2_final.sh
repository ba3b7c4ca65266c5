# end-of-round validation: the GPU test suite, smoke(), the default bench line and the reference arm (run under gpurun)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
(time timeout 900 python bench.py) > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/r2_final_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_final_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'points_per_s', 'gpu_launches')})
print('e2e', d['e2e']['ms_per_step'], 'pageable', d['e2e_pageable']['ms_per_step'], 'roofline', d['roofline']['frac'], d['roofline']['kernel_ms'])
x = d.get('extra', {})
for k in x:
    if k not in ('fft_fr', 'das_fft_extension', 'msm_2p24', 'adversarial_blobs', 'adversarial_msm', 'variable_base_e2e'):
        print(k, json.dumps(x[k])[:700])
print('fft_fr', {k: round(v['ms'] * 1e3, 1) for k, v in x['fft_fr'].items()})
print('adv', x['adversarial_msm']['worst_vs_uniform'], x['adversarial_blobs']['worst_vs_uniform'])
PY
(time timeout 600 python bench.py --impl reference --steps 3 --warmup 3) > gpurun_out/r2_final_ref.json 2>&1
tail -c 400 gpurun_out/r2_final_ref.json
