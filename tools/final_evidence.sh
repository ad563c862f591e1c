#!/bin/bash
# usage (after the GPU capture below has been brought back in gpurun_out/):  tools/final_evidence.sh TAG
# Regenerates the judged summaries under profiles/ from gpurun_out/TAG_launches.csv and gpurun_out/prof_TAG_all.ncu-rep.
# GPU side (one gpurun call, each command after the previous one exited):
#   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-exact
#   ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/TAG_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-exact
#   ncu --set full --clock-control none --import-source on -k regex:"hessian_eigen|gauss_z_tma|gauss_xy" -s 12 -c 12 -o gpurun_out/prof_TAG_all python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-exact
set -e
TAG=$1
cd "$(dirname "$0")/.."
REP=gpurun_out/prof_${TAG}_all.ncu-rep
cp gpurun_out/${TAG}_launches.csv profiles/${TAG}_launches.csv
python tools/launch_summary.py profiles/${TAG}_launches.csv > profiles/${TAG}_launches_summary.txt
( echo "# ncu --set full --clock-control none --import-source on -k regex:hessian_eigen|gauss_z_tma|gauss_xy -s 12 -c 12"
  echo "# command: python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-exact   (2048x2048x512, sigma 2,4,6, fma smoothing; build of $(git rev-parse --short HEAD))"
  python tools/ncu_summary.py $REP ) > profiles/${TAG}_ncu_summary.txt
# kernel ids inside the report: the 12 captured launches of one step in stream order
IDS=$(ncu -i $REP --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin)); h = rows[0]
a = c = None
for r in rows[2:]:
    n = r[h.index('Kernel Name')]
    if 'hessian_eigen_kernel' in n and a is None: a = r[h.index('ID')]
    if 'hessian_eigen_compact' in n: c = r[h.index('ID')]
print(int(a) + 1, int(c) + 1)   # --kernel-id counts invocations from 1")
python tools/k3_executed.py $REP $IDS > profiles/${TAG}_k3_executed.txt
for k in hessian_eigen_compact_kernel "hessian_eigen_kernelILi0ELb0" hessian_eigen_shell_kernelILi0 gauss_xy_warp_kernelILi6 "gauss_z_tma_kernelILi9ELb0"; do
  echo "== $k"
  cuobjdump -sass pnr_b200/_lib/libfrangi_gpu.so | awk -v k="$k" '/Function :/{f=index($0,k)>0} f{print}' | grep -E "^\s+/\*[0-9a-f]{4}\*/" \
    | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?//' | awk '{print $1}' | sed 's/\..*//' | sort | uniq -c | sort -rn | awk '{printf "%s:%s ", $2, $1} END{print ""}'
done > profiles/${TAG}_sass_static.txt
# measured DRAM bytes of the K3 class per launch (three tile launches + three shell launches of one step, / 3)
ncu -i $REP --page raw --csv 2>/dev/null | python -c "
import csv, json, sys
rows = list(csv.reader(sys.stdin)); h = rows[0]; u = rows[1]
tot = 0.0
for r in rows[2:]:
    if 'hessian' not in r[h.index('Kernel Name')]: continue
    for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        v = float(r[h.index(m)]); unit = u[h.index(m)]
        tot += v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}[unit]
p = 'profiles/traffic.json'
d = json.load(open(p))
d['hessian_eigen:2048x2048x512'] = int(tot / 3)
d['_source'] = 'profiles/${TAG}_ncu_summary.txt (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum of the three tile launches and three shell launches of one step, divided by 3): per launch of the class, like roofline.achieved'
json.dump(d, open(p, 'w'), indent=1)
print('K3 class DRAM bytes per launch:', tot / 3)"
