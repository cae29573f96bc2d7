#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 > gpurun_out/r02e_gputests.log; tail -6 gpurun_out/r02e_gputests.log
python tools/cells_bench.py > gpurun_out/r02e_cells.log 2>&1; cat gpurun_out/r02e_cells.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02e_bench.err
export OFC_FIT_ITERS=4
F="python tools/fit_profile.py f32_1Mx128_k1024"
$F > gpurun_out/r02e_fit_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02e_fit_launches.csv $F > gpurun_out/r02e_fit_ncu.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/r02e_fit_plain.log
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(l for l in open("gpurun_out/r02e_fit_launches.csv") if l.startswith(chr(34)))]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
agg=collections.OrderedDict()
for r in rows[1:]:
    v=float(r[vi].replace(",","")); u=r[ui]
    v*= {"ns":1e-3,"us":1.0,"ms":1e3,"nsecond":1e-3,"usecond":1.0,"msecond":1e3}.get(u,1.0)
    a=agg.setdefault(r[ki][:80],[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]: print(f"  {n:4d} {t:12.1f} us  {k}")
PY
