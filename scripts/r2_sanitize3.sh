# compute-sanitizer over what changed after r2_sanitize2.sh: k_ntt_cluster (DSMEM), k_g1_stage3_* (fused triples), the new
# k_group_finish, k_direct_msm with Jacobian output behind prepare_msm, the one-pass cells + proofs with coalesced callers
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  for mode in none cells widemsm round2; do
    arg=$mode; [ $mode = none ] && arg=""
    timeout 1700 compute-sanitizer --tool $tool python scripts/sanitize_target.py $arg > gpurun_out/r2_sanitize3_${tool}_${mode}.log 2>&1
    echo "$tool $mode rc=$?: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize target ok' gpurun_out/r2_sanitize3_${tool}_${mode}.log | tr '\n' ' ')"
  done
done
