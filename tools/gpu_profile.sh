#!/bin/bash
# ncu evidence for the round: launch list of one profiled run + full-set captures of the top kernels.
mkdir -p gpurun_out
KREG='regex:prefilter|polyexp|flow_iter|bgr2gray|flow_encode|grid_cells|minmax'
python tools/profile_step.py > gpurun_out/profile_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 200 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:flow_iter -s 33 -c 1 -o gpurun_out/prof_flow_iter -f python tools/profile_step.py > gpurun_out/ncu_iter.log 2>&1; echo "iter rc=$?"
ncu --set full --clock-control none --import-source on -k regex:polyexp -s 11 -c 1 -o gpurun_out/prof_polyexp -f python tools/profile_step.py > gpurun_out/ncu_poly.log 2>&1; echo "poly rc=$?"
ncu --set full --clock-control none --import-source on -k regex:prefilter -s 9 -c 3 -o gpurun_out/prof_prefilter -f python tools/profile_step.py > gpurun_out/ncu_pref.log 2>&1; echo "pref rc=$?"
ls -la gpurun_out
