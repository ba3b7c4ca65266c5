cd $GRAFT_REPO_ROOT
for m in 0 3 2; do B200_ACC_KARA=$m python scripts/acc_time.py 20 kara$m | cut -c1-160; done
