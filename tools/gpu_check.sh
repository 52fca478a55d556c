#!/bin/bash
# Mid-round check on one GPU: the whole -m gpu suite and a short bench (no CPU baseline).
# Usage: gpurun --timeout 1200 -- 'timeout 1150 bash tools/gpu_check.sh r02c'
TAG=${1:-chk}
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/${TAG}_pytest.log | tail -8
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"
python - "$TAG" <<'PY'
import json,sys
tag=sys.argv[1]
try:
    d=json.loads([l for l in open("gpurun_out/%s_bench.json"%tag) if l.startswith("{")][0])
    print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],"objf_check",d.get("objf_check",{}).get("rel_diff"),"ctc",d["ctc_roofline"]["frac"],d["ctc_roofline"]["ms_per_call"],"gemm",d["gemm_roofline"].get("frac"), d["gemm_roofline"].get("achieved"),"fs3",d["fs3"]["ms_per_step"],"gru",d["configs3_gru"]["ms_per_step"])
    print(d["roofline"].get("ms_per_step_by_kernel"))
except Exception as e: print("ERR",e)
PY
