cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_msm.py -m gpu -x -q -k "baseline or large or adversarial" 2>&1 | tail -3
for v in 1 0; do B200_MSM_PAGEABLE_STAGING=$v python bench.py --no-extra 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('staging $v', 'device %.3f' % d['ms_per_step'], 'e2e pinned %.3f' % d['e2e']['ms_per_step'], 'pageable %.3f' % d['e2e_pageable']['ms_per_step'])"; done
