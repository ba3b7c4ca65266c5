#!/bin/bash
# round 2: regenerates the ncu launch lists and one --set full capture of k_accumulate under gpurun_out/ (run under gpurun);
# tables land in gpurun_out/r02_launches.md
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
run() { # name first-kernel target...
  local name=$1 first=$2; shift 2
  $NCU -c 3000 --log-file gpurun_out/r02_launches_$name.csv python scripts/ncu_target.py "$@" > /dev/null 2>&1
  echo "## $name -- \`python scripts/ncu_target.py $*\`" >> gpurun_out/r02_launches.md
  echo >> gpurun_out/r02_launches.md
  python scripts/launch_table.py gpurun_out/r02_launches_$name.csv "$first" >> gpurun_out/r02_launches.md
  echo >> gpurun_out/r02_launches.md
}
: > gpurun_out/r02_launches.md
run msm_2p20 "k_digits<0>" msm 20 2
B200_MSM_DIRECT=0 run msm_2p12_buckets "k_digits<0>" msm 12 2
run msm_2p12_direct "k_fr_from_mont" msm 12 2
run fk20_1 "k_blob_to_fr" fk20 1 1
run blob64_commit "k_blob_to_fr" blob 64 2
run blob1_commit "k_blob_to_fr" blob 1 2
run blob8_commit "k_blob_to_fr" blob 8 2
run proof64 "k_blob_to_fr" proof 64 2
run fk20_64 "k_blob_to_fr" fk20 64 1
run ntt_2p20 "k_ntt_pass" ntt 20 3
run ntt_2p12 "k_ntt_cluster" ntt 12 3
ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 1 -c 1 -f -o /tmp/prof_acc python scripts/ncu_target.py msm 20 2 > /dev/null 2>&1
ncu -i /tmp/prof_acc.ncu-rep --page raw --csv > gpurun_out/r02_accumulate_raw.csv
python scripts/ncu_raw_table.py gpurun_out/r02_accumulate_raw.csv > gpurun_out/r02_accumulate_table.md
ncu --set full --clock-control none -k regex:k_direct_msm -s 1 -c 1 -f -o /tmp/prof_direct python scripts/ncu_target.py blob 64 2 > /dev/null 2>&1
ncu -i /tmp/prof_direct.ncu-rep --page raw --csv > gpurun_out/r02_direct_msm_raw.csv
python scripts/ncu_raw_table.py gpurun_out/r02_direct_msm_raw.csv > gpurun_out/r02_direct_msm_table.md
ncu --set full --clock-control none -k regex:k_direct_lincomb -s 1 -c 1 -f -o /tmp/prof_lincomb python scripts/ncu_target.py fk20 64 1 > /dev/null 2>&1
ncu -i /tmp/prof_lincomb.ncu-rep --page raw --csv > gpurun_out/r02_direct_lincomb_raw.csv
python scripts/ncu_raw_table.py gpurun_out/r02_direct_lincomb_raw.csv > gpurun_out/r02_direct_lincomb_table.md
cat gpurun_out/r02_launches.md | head -80
cat gpurun_out/r02_accumulate_table.md
cat gpurun_out/r02_direct_msm_table.md
