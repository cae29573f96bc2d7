#!/bin/bash
# ncu evidence for the tensor-core k-means path: launch list + full-set capture of the E-step / M-step kernels
# on the 1M x 512, k = 256 float32 shape (tools/tc_check.py one ...).  One ncu "use" per gpurun call.
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/launches.csv
export PYTHONPATH=.
CMD="python tools/tc_check.py one 1000000 512 256"
$CMD > gpurun_out/tc_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
for k in kmeans_assign_tc seg_sums; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 2 -o gpurun_out/prof_$k -f $CMD > gpurun_out/ncu_$k.log 2>&1; echo "$k rc=$?"
done
ls -la gpurun_out | tail -6
