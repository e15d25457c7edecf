#!/usr/bin/env python
"""One region's kernels out of an ncu launch list (the --metrics gpu__time_duration.sum,smsp__inst_executed.sum,
dram__bytes_read.sum,dram__bytes_write.sum pass of tools/gpu_iter.sh): duration, warp instructions and DRAM bytes per
launch, from the second k_read_prep to the third.  Usage: launch_list.py launches.csv"""
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
for i,r in enumerate(rows):
    if r and r[0]=="ID": h=i;break
hdr=rows[h]
ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value")
from collections import OrderedDict
d=OrderedDict()
for r in rows[h+1:]:
    d.setdefault((int(r[0]),r[ki].split('(')[0]),{})[r[mi]]=r[vi]
# find second k_read_prep and print until third
starts=[i for (i,k) in d if k=='k_read_prep']
tot=0
for (i,k),m in d.items():
    if starts[1]<=i<starts[2]:
        t=float(m['gpu__time_duration.sum'])/1e3; tot+=t
        print(i,k, "%.1f us"%t, "%.1fM inst"%(float(m['smsp__inst_executed.sum'])/1e6), "rd %.0f MB wr %.0f MB"%(float(m['dram__bytes_read.sum'])/1e6, float(m['dram__bytes_write.sum'])/1e6))
print("total %.1f us"%tot)
