#!/bin/bash
# round-2 GPU call: parity tests, per-cell k-means M-step variants, bench variants, ncu capture of the flow kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02c_gputests.log; tail -4 gpurun_out/r02c_gputests.log
for m in 0 1 2; do
  OFC_CELLS_MSTEP=$m python tools/cells_bench.py quick > gpurun_out/r02c_cells_m$m.log 2>&1; echo "mstep=$m"; cat gpurun_out/r02c_cells_m$m.log
done
B="python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline"
$B > gpurun_out/r02c_bench_default.json 2> gpurun_out/r02c_bench.err; echo "default rc=$?"
OFC_FUSE_UPSAMPLE=0 $B > gpurun_out/r02c_bench_nofuseups.json 2>> gpurun_out/r02c_bench.err
OFC_PREFILTER_PYR=0 $B > gpurun_out/r02c_bench_nopyr.json 2>> gpurun_out/r02c_bench.err
OFC_FUSE_GRID=0 $B > gpurun_out/r02c_bench_nofusegrid.json 2>> gpurun_out/r02c_bench.err
python - <<'PY'
import json
for n in ("default","nofuseups","nopyr","nofusegrid"):
    try:
        b=json.load(open(f"gpurun_out/r02c_bench_{n}.json"))
        print(n, round(b["value"],1), round(b["ms_per_step"],3), {k:round(v["ms_per_step"],3) for k,v in b["kernels"].items()})
    except Exception as e:
        print(n, "failed", e)
PY
export OFC_CHUNK=33
P="python tools/profile_step.py"
$P > gpurun_out/r02c_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'flow_iter_tmem|prefilter_pyr|flow_encode_grid|polyexp_strip' -s 24 -c 12 -o gpurun_out/r02c_flow -f $P > gpurun_out/r02c_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02c_ncu.log; ls -la gpurun_out/*.ncu-rep
