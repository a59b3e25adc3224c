#!/bin/bash
# DRAM bytes and duration per GEMM launch (gate_up, ff_out, qkv, attn_out at config-2 shapes) under ncu for rasterisation
# settings "group serpentine by_n" (EXPERIMENTS build)
mkdir -p gpurun_out
export MMADA_B200_LIB=$PWD/mmada_b200/libmmada_b200_exp.so
for cfg in "${@:-0 1 0}"; do
  set -- $cfg
  export MMADA_GEMM_GROUP_M=$1 MMADA_GEMM_SERPENTINE=$2 MMADA_GEMM_BY_N=$3
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_kernel -s 4 -c 4 --csv \
      python scripts/gemm_traffic_probe.py 2>/dev/null | python -c "
import csv,sys
rows=[r for r in csv.reader(sys.stdin) if len(r)>10 and r[0].isdigit()]
d={}
for r in rows:
    v=float(r[-1].replace(',',''))
    u=r[-2]
    if 'byte' in u.lower():
        v*= {'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}.get(u,1)
    d.setdefault(r[0],{})[r[-3]]=v
print('group=$1 serpentine=$2 by_n=$3:', '  '.join('%.2f GB %.0f us' % ((v.get('dram__bytes_read.sum',0)+v.get('dram__bytes_write.sum',0))/1e9, v.get('gpu__time_duration.sum',0)/ (1e3 if v.get('gpu__time_duration.sum',0)>1e5 else 1)) for v in d.values()))
"
done
