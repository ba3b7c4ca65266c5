set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_run1_smi.txt 2>&1
(time timeout 1500 python -m pytest tests -m gpu -x -q --durations=25) > gpurun_out/r2_run1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_run1_pytest.log
(time timeout 900 python bench.py --steps 20 --warmup 5) > gpurun_out/r2_run1_bench.json 2> gpurun_out/r2_run1_bench.err
echo "bench rc=$?" >> gpurun_out/r2_run1_bench.err
tail -5 gpurun_out/r2_run1_pytest.log
tail -c 1500 gpurun_out/r2_run1_bench.json
