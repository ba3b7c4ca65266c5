#!/bin/bash
# regenerates the ncu launch lists (and one --set full capture of k_accumulate) under gpurun_out/; run under gpurun
set -x
mkdir -p gpurun_out
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
$NCU -c 400 --log-file gpurun_out/launches_msm_20.csv python scripts/ncu_target.py msm 20 2 > /dev/null 2>&1
$NCU -c 400 --log-file gpurun_out/launches_msm_12.csv python scripts/ncu_target.py msm 12 2 > /dev/null 2>&1
$NCU -c 400 --log-file gpurun_out/launches_blob64.csv python scripts/ncu_target.py blob 64 2 > /dev/null 2>&1
$NCU -c 400 --log-file gpurun_out/launches_blob1.csv python scripts/ncu_target.py blob 1 2 > /dev/null 2>&1
$NCU -c 600 --log-file gpurun_out/launches_proof64.csv python scripts/ncu_target.py proof 64 2 > /dev/null 2>&1
$NCU -c 400 --log-file gpurun_out/launches_ntt20.csv python scripts/ncu_target.py ntt 20 3 > /dev/null 2>&1
$NCU -c 3000 --log-file gpurun_out/launches_fk20.csv python scripts/ncu_target.py fk20 16 1 > /dev/null 2>&1
$NCU -c 600 --log-file gpurun_out/launches_verify64.csv python scripts/ncu_target.py verify 64 2 > /dev/null 2>&1
$NCU -c 600 --log-file gpurun_out/launches_verifycells.csv python scripts/ncu_target.py verifycells 128 2 > /dev/null 2>&1
$NCU -c 600 --log-file gpurun_out/launches_recover.csv python scripts/ncu_target.py recover 64 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pairing_check -s 1 -c 1 -f -o /tmp/prof_pairing \
    python scripts/ncu_target.py verify 64 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 1 -c 1 -f -o /tmp/prof_accumulate_2p20 \
    python scripts/ncu_target.py msm 20 2 > /dev/null 2>&1
ls -la gpurun_out
# the .ncu-rep files stay on the box (gpurun_out/ is capped at 64 MiB): export the raw pages as CSV instead
ncu -i /tmp/prof_pairing.ncu-rep --page raw --csv > gpurun_out/prof_pairing_raw.csv
ncu -i /tmp/prof_accumulate_2p20.ncu-rep --page raw --csv > gpurun_out/prof_accumulate_2p20_raw.csv
ncu --set full --clock-control none -k regex:k_segment_fold -s 1 -c 1 -f -o /tmp/prof_segment_fold python scripts/ncu_target.py msm 20 2 > /dev/null 2>&1
ncu -i /tmp/prof_segment_fold.ncu-rep --page raw --csv > gpurun_out/prof_segment_fold_raw.csv
ls -la gpurun_out
