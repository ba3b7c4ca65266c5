cd $GRAFT_REPO_ROOT
lscpu | grep -i -E "numa|socket|model name|^cpu\(s\)" 
ls /sys/devices/system/node/ 2>/dev/null | head
for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q "^0x0302" $d/class; then echo $(basename $d) numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist); fi; done
nvidia-smi topo -m 2>&1 | head -20
python - <<'PY'
import os
print('affinity', len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:4], '...')
print('cpu_count', os.cpu_count())
PY
cat /sys/devices/system/node/node*/cpulist 2>/dev/null
free -g | head -2
