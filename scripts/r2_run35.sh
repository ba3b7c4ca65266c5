cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1400 python -m pytest tests/test_gpu_msm.py tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -3
B200_BLOB_DIRECT=0 B200_FK20_DIRECT=0 timeout 900 python -m pytest tests/test_gpu_eip4844.py -m gpu -x -q -k "not direct and not fk20_lincomb" 2>&1 | tail -3
python bench.py --no-extra 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('device %.3f' % d['ms_per_step'], 'e2e pinned %.3f' % d['e2e']['ms_per_step'], 'roofline', d['roofline']['frac'], d['roofline']['kernel_ms'])"
ncu --metrics gpu__time_duration.sum --clock-control none --csv -c 3000 --log-file gpurun_out/r2_l20.csv python scripts/ncu_target.py msm 20 2 > /dev/null 2>&1
python scripts/launch_table.py gpurun_out/r2_l20.csv "k_digits<0>" | tail -8
timeout 600 compute-sanitizer --tool racecheck python scripts/sanitize_target.py widemsm 2>&1 | grep -E "RACECHECK SUMMARY|sanitize target ok|Error" | head -5
