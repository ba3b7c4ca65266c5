"""Turn an ncu `--metrics gpu__time_duration.sum --csv` log into a markdown launch table (last repetition only).

usage: python scripts/launch_table.py gpurun_out/launches_x.csv [first_kernel_substring]
The table starts at the LAST launch whose name contains first_kernel_substring (default: k_digits<0>, the first
kernel of one MSM), so warm-up repetitions and one-time preparation kernels are dropped."""
import csv
import re
import sys

path, first = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"^(void )?(b200::)?", "", r["Kernel Name"])
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"<\(bool\)([01])>", r"<\1>", name)
    rows.append((name, r["Grid Size"], r["Block Size"], float(r["Metric Value"].replace(",", "")) / 1e3))
start = 0
if first:
    # start of the last complete repetition: the last launch of `first` that is preceded by a different kernel
    idx = [i for i, r in enumerate(rows) if first in r[0] and (i == 0 or first not in rows[i - 1][0])]
    start = idx[-1] if idx else 0
rows = rows[start:]
total = sum(r[3] for r in rows)
print("| kernel | grid | block | us | share |\n|---|---|---|---:|---:|")
# merge consecutive identical launches
i = 0
while i < len(rows):
    j = i
    while j + 1 < len(rows) and rows[j + 1][:3] == rows[i][:3]:
        j += 1
    us = sum(r[3] for r in rows[i:j + 1])
    cnt = j - i + 1
    nm = rows[i][0] + (f" x{cnt}" if cnt > 1 else "")
    print(f"| {nm} | {rows[i][1]} | {rows[i][2]} | {us:.1f} | {100 * us / total:.1f}% |")
    i = j + 1
print(f"| **total** | | | **{total:.1f}** | |")
