#!/bin/bash
# full GPU check: all gpu tests, smoke, bench (usage: gpu_full.sh TAG [bench args])
TAG=${1:-full}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_all_tests.log 2>&1; echo "all tests rc=$?" >> gpurun_out/${TAG}_all_tests.log; tail -6 gpurun_out/${TAG}_all_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 1200 python bench.py "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench.json"))
    for k in ("value","ms_per_step","clocks","single_stream_ms_per_frame","given_order"):
        print(k, d.get(k))
    print("e2e", {k:v for k,v in d["e2e"].items() if k!="note"})
    print("roofline", {k:v for k,v in (d.get("roofline") or {}).items() if k not in ("how","longest_kernel_of_frame")})
    print("stages", json.dumps({k:{kk:vv for kk,vv in v.items() if kk in ("us","frac","sum_of_kernel_us")} for k,v in d.get("stages",{}).items()}))
    print("kernels", {k:round(v["us_per_frame"],1) for k,v in d.get("kernels",{}).items()})
    w=d.get("workloads",{})
    print("workloads", json.dumps({k:{kk:(round(vv,1) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ("us","pillars","frac")} for k,v in w.items() if k!="NMS20k_grid"}))
    g=w.get("NMS20k_grid",{})
    print("nms grid entries", len(g), {k:g[k] for k in list(g)[:4]})
    for k in ("e2e_dropin","naive_gpu","cpu_baseline"):
        print(k, {kk:vv for kk,vv in (d.get(k) or {}).items() if kk not in ("note","sample","calibration")})
except Exception as e:
    print("bench parse failed", e)
PY
