#!/bin/bash
# bf16 GEMM operands, clip/self-repair ABI (no test yet), full GPU suite, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x -s > gpurun_out/c7_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed" gpurun_out/c7_pytest.log | tail -3
timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err
echo "bench rc=$?"
B200RNN_NO_BF16_GEMM=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ctc-roofline --no-objf-check > gpurun_out/c7_bench_nobf16.json 2> gpurun_out/c7_bench_nobf16.err
python - <<'PY'
import json
for f in ("gpurun_out/c7_bench.json","gpurun_out/c7_bench_nobf16.json"):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][0])
        print(f, d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("objf_check"), d["roofline"]["ms_per_step_by_kernel"])
    except Exception as e: print(f, "ERR", e)
PY
