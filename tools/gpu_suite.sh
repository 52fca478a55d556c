#!/bin/bash
# One GPU-box call: the whole -m gpu suite, the bench (both arms), and the ncu evidence for profiles/.
# Every step runs under its own timeout so that a hung kernel costs seconds, not the call.
# Usage (from the repo root):  gpurun --timeout 1500 -- 'timeout 1450 bash tools/gpu_suite.sh r02'
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed" gpurun_out/${TAG}_pytest.log | tail -2
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
echo "bench rc=$?"
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
echo "bench ref rc=$?"
# ncu: launch list of the same bench command (training steps only), then one full capture of the dominant kernels
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-objf-check --no-ctc-roofline > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rec_tc -s 10 -c 2 -o gpurun_out/${TAG}_prof_rec -f \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-objf-check --no-ctc-roofline > gpurun_out/${TAG}_ncu_rec.log 2>&1
echo "ncu rec rc=$?"
python - <<'PY'
import json,sys
tag=sys.argv[1] if len(sys.argv)>1 else "r02"
try:
    d=json.loads([l for l in open("gpurun_out/%s_bench_n1.json"%tag) if l.startswith("{")][0])
    print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],"objf_check",d["objf_check"]["rel_diff"],"ctc",d["ctc_roofline"]["frac"],d["ctc_roofline"]["ms_per_call"],"gemm",d["gemm_roofline"]["frac"],"fs3",d["fs3"]["ms_per_step"],"gru",d["configs3_gru"]["ms_per_step"],"cpu",d["cpu_baseline"]["value"])
except Exception as e: print("ERR",e)
PY
