cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for mode in 0 1; do
B200_MSM_AFFINE_TREE=$mode ncu --set full --clock-control none --import-source on -k regex:k_accumulate_affine -s 1 -c 1 -f -o /tmp/prof_aff_$mode \
    python scripts/ncu_target.py msm 20 2 > /dev/null 2>&1
ncu -i /tmp/prof_aff_$mode.ncu-rep --page raw --csv > gpurun_out/r2_prof_affine_tree${mode}_raw.csv
python scripts/ncu_raw_table.py gpurun_out/r2_prof_affine_tree${mode}_raw.csv
ncu -i /tmp/prof_aff_$mode.ncu-rep --page source --csv > /tmp/src_$mode.csv 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open('/tmp/src_$mode.csv')))
hdr=rows[0]
print(hdr[:12])
# aggregate stall samples per source line
try:
    si=hdr.index('# Samples') if '# Samples' in hdr else None
except Exception: si=None
print(len(rows))
PY
head -c 3000 /tmp/src_$mode.csv > gpurun_out/r2_src_head_$mode.txt
done
