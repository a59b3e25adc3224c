#!/bin/bash
# DRAM bytes and duration per GEMM launch under ncu for several rasterisation / cache-hint settings
mkdir -p gpurun_out
for cfg in "16 ln" "8 ln" "4 ln" "8 nn" "12 ln" "6 ll" "24 ln" "8 lf"; do
  set -- $cfg
  export MMADA_GEMM_GROUP_M=$1 MMADA_GEMM_HINTS=$2
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_kernel -s 4 -c 4 --csv \
      python scripts/gemm_traffic_probe.py 2>/dev/null | python -c "
import csv,sys
rows=[r for r in csv.reader(sys.stdin) if len(r)>10 and r[0].isdigit()]
d={}
for r in rows: d.setdefault(r[0],{})[r[-3]]=float(r[-1].replace(',',''))
print('group_m=$1 hints=$2:', '  '.join('%.2f GB %.0f us' % ((v.get('dram__bytes_read.sum',0)+v.get('dram__bytes_write.sum',0))/(1e9 if max(v.values())>1e6 else 1), v.get('gpu__time_duration.sum',0)) for v in d.values()))
"
done
