# compute-sanitizer over the final round-2 kernels (13-bit direct tables, balanced k_direct_msm, one-pass cells + proofs)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  for mode in cells round2; do
    ( time timeout 1700 compute-sanitizer --tool $tool python scripts/sanitize_target.py $mode > gpurun_out/r2_sanitize2_${tool}_${mode}.log 2>&1 ) 2> gpurun_out/r2_sanitize2_time.txt
    echo "$tool $mode rc=$?: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize target ok' gpurun_out/r2_sanitize2_${tool}_${mode}.log | tr '\n' ' ') $(grep real gpurun_out/r2_sanitize2_time.txt)"
  done
done
B200_BLOB_DIRECT=16 B200_BLOB_DIRECT_BITS=8 B200_FK20_DIRECT_BITS=8 timeout 1500 compute-sanitizer --tool memcheck python scripts/sanitize_target.py round2 > gpurun_out/r2_sanitize2_memcheck_round2_bits8.log 2>&1
echo "memcheck round2 (8-bit tables, bucket engine above 16 blobs) rc=$?: $(grep -E 'ERROR SUMMARY|sanitize target ok' gpurun_out/r2_sanitize2_memcheck_round2_bits8.log | tr '\n' ' ')"
