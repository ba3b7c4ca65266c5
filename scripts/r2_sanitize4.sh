# compute-sanitizer over the last additions: k_g1_stageR_* (six fused stages), split scalar multiplications, team-form FK20
# lincombs (k_direct_msm with per-vector point blocks), split verification lincombs
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1700 compute-sanitizer --tool $tool python scripts/sanitize_target.py cells > gpurun_out/r2_sanitize4_${tool}_cells.log 2>&1
  echo "$tool cells rc=$?: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize target ok' gpurun_out/r2_sanitize4_${tool}_cells.log | tr '\n' ' ')"
done
B200_FFT_G1_FUSE=3 B200_FFT_G1_SPLIT=1 timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_target.py cells > gpurun_out/r2_sanitize4_racecheck_cells_f3.log 2>&1
echo "racecheck cells (split triples) rc=$?: $(grep -E 'RACECHECK SUMMARY|sanitize target ok' gpurun_out/r2_sanitize4_racecheck_cells_f3.log | tr '\n' ' ')"
