#!/bin/bash
# tensor-core k-means check + flow-flag tests + ncu launch list of one shape
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 300 python tools/tc_check.py time > gpurun_out/tc_check.log 2>&1; echo "tc_check rc=$?"
tail -32 gpurun_out/tc_check.log
timeout 600 python -m pytest tests/test_gpu_flow.py -x -q -k "flags" > gpurun_out/pytest_flags.log 2>&1; echo "pytest flags rc=$?"
tail -5 gpurun_out/pytest_flags.log
for shape in "1000000 64 256" "1000000 512 256"; do
  tag=$(echo $shape | tr ' ' '_')
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/tc_launches_$tag.csv python tools/tc_check.py one $shape > gpurun_out/tc_ncu_$tag.log 2>&1
  echo "ncu $tag rc=$?"
done
python - <<PY
import csv,collections
for tag in ("1000000_64_256","1000000_512_256"):
    rows=[r for r in csv.reader(l for l in open(f"gpurun_out/tc_launches_{tag}.csv") if l.startswith(chr(34)))]
    h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
    agg=collections.OrderedDict()
    for r in rows[1:]:
        v=float(r[vi].replace(",","")); u=r[ui]
        v*= {"ns":1e-3,"us":1.0,"ms":1e3,"nsecond":1e-3,"usecond":1.0,"msecond":1e3}.get(u,1.0)
        a=agg.setdefault(r[ki][:70],[0,0.0]); a[0]+=1; a[1]+=v
    print(tag)
    for k,(n,t) in agg.items(): print(f"  {n:4d} {t:10.1f} us  {k}")
PY
