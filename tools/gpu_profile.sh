#!/bin/bash
# ncu evidence: launch list of one profiled run + full-set captures of the top kernels.
python -c "from opticalflowclustering_b200 import _build; _build.build()"
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
KREG='regex:prefilter|polyexp|flow_iter|flow_upsample|bgr2gray|flow_encode|grid_cells|minmax'
python tools/profile_step.py > gpurun_out/profile_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 200 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
# third step of the script: skip the launches of the first two steps (launches per step read from the list)
ncu --set full --clock-control none --import-source on -k regex:flow_iter_strip -s 33 -c 1 -o gpurun_out/prof_flow_iter_strip -f python tools/profile_step.py > gpurun_out/ncu_iter.log 2>&1; echo "iter rc=$?"
ncu --set full --clock-control none --import-source on -k regex:polyexp_strip -s 11 -c 1 -o gpurun_out/prof_polyexp_strip -f python tools/profile_step.py > gpurun_out/ncu_poly.log 2>&1; echo "poly rc=$?"
ncu --set full --clock-control none --import-source on -k regex:prefilter_direct -s 6 -c 3 -o gpurun_out/prof_prefilter_direct -f python tools/profile_step.py > gpurun_out/ncu_pref.log 2>&1; echo "pref rc=$?"
ls -la gpurun_out | tail -12
