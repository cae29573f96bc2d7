#!/bin/bash
# ncu evidence for the k-means kernels: launch list + full-set capture of the tensor-core E-step, the CSR sums
# and the fused uint8 step.  One ncu "use" per gpurun call.
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/launches.csv
export PYTHONPATH=.
CMD="python tools/tc_check.py one 1000000 128 1024"
$CMD > gpurun_out/tc_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
for k in kmeans_assign_tc seg_sums; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 2 -o gpurun_out/prof_$k -f $CMD > gpurun_out/ncu_$k.log 2>&1; echo "$k rc=$?"
done
CMD2="python tools/step_check.py"
$CMD2 > gpurun_out/step_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:kmeans_step_u8\|kmeans_assign_medium -s 2 -c 3 -o gpurun_out/prof_kmeans_step -f $CMD2 > gpurun_out/ncu_step.log 2>&1; echo "step rc=$?"
tail -3 gpurun_out/step_plain.log
ls -la gpurun_out | tail -6
