#!/bin/bash
# Final evidence of a round on ONE B200: parity tests, smoke, bench (both arms), ncu launch list of the bench step,
# ncu --set full of the dominant kernels.  Usage: bash tools/gpu_final.sh <tag>   (writes gpurun_out/<tag>_*)
TAG=${1:-r02h}
export PYTHONPATH=.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.csv 2>&1
python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_gputests_full.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_gputests_full.log
grep -E "labels differ|flow parity" gpurun_out/${TAG}_gputests_full.log > gpurun_out/${TAG}_parity_lines.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; echo "reference rc=$?"
SHORT="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline"
$SHORT > gpurun_out/${TAG}_plain.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $SHORT > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "ncu list rc=$?"
export OFC_CHUNK=33
P="python tools/profile_step.py"
$P > gpurun_out/${TAG}_plain2.log 2>&1 && ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'flow_iter|prefilter|flow_encode|polyexp|flow_upsample|bgr2gray' -c 24 -o gpurun_out/${TAG}_flow -f $P > gpurun_out/${TAG}_ncu_flow.log 2>&1
echo "ncu flow rc=$?"
export OFC_CHUNK=9 OFC_K=8
$P > gpurun_out/${TAG}_plain3.log 2>&1 && ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'kmeans_cells_fast' -c 1 -o gpurun_out/${TAG}_cells -f $P > gpurun_out/${TAG}_ncu_cells.log 2>&1
echo "ncu cells rc=$?"
unset OFC_K
S="python tools/step_check.py"
$S > gpurun_out/${TAG}_plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'kmeans_step_u8d4' -s 2 -c 2 -o gpurun_out/${TAG}_step -f $S > gpurun_out/${TAG}_ncu_step.log 2>&1
echo "ncu step rc=$?"
# gpurun returns at most 64 MiB: the raw pages travel as CSV, only the flow report itself is kept (source view of the top kernel)
for n in flow cells step; do
  ncu -i gpurun_out/${TAG}_${n}.ncu-rep --page raw --csv > gpurun_out/${TAG}_${n}_raw.csv 2>/dev/null
done
ncu -i gpurun_out/${TAG}_flow.ncu-rep --page source --csv --kernel-name regex:flow_iter_tmem --launch-count 1 > gpurun_out/${TAG}_flow_iter_source.csv 2>/dev/null
rm -f gpurun_out/${TAG}_cells.ncu-rep gpurun_out/${TAG}_step.ncu-rep
ls -la gpurun_out/ | tail -30; du -sm gpurun_out
