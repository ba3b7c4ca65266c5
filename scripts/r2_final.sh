cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
(time timeout 900 python bench.py --steps 20 --warmup 5) > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/r2_final_bench.err
(time timeout 600 python bench.py --impl reference --steps 5 --warmup 3) > gpurun_out/r2_final_ref.json 2>&1
tail -c 600 gpurun_out/r2_final_ref.json
