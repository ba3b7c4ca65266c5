cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  for mode in round2 widemsm cells; do
    timeout 1500 compute-sanitizer --tool $tool python scripts/sanitize_target.py $mode > gpurun_out/r2_sanitize_${tool}_${mode}.log 2>&1
    echo "$tool $mode rc=$?: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize target ok' gpurun_out/r2_sanitize_${tool}_${mode}.log | tr '\n' ' ')"
  done
done
B200_ACC_KARA=3 timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_target.py widemsm > gpurun_out/r2_sanitize_memcheck_kara.log 2>&1
echo "memcheck kara rc=$?: $(grep -E 'ERROR SUMMARY|sanitize target ok' gpurun_out/r2_sanitize_memcheck_kara.log | tr '\n' ' ')"
