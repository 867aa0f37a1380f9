#!/bin/bash
# per-kernel durations of one eager forward of a CNN config (ncu launch list, forward kernels only)
set -u
O=gpurun_out
c=${1:-resnet18_t2a8}
QTB200_BENCH_GRAPH=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tc_gemm|image_|pool_|act_quant|rowsum|gemm|elementwise|reduce' -s ${2:-400} -c ${3:-80} --csv --log-file $O/l_$c.csv python bench.py --config $c --steps 1 --warmup 3 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/l_$c.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
for r in rows[1:]:
    print(r[ki][:90], r[vi])
PY
