#!/bin/bash
# End-of-round evidence on ONE GPU: the whole -m gpu suite, the bench (both arms), the ncu launch list of the bench
# command and one `ncu --set full` capture per kernel family (each only after its command has exited 0 without ncu).
# Usage (from the repo root):  gpurun --timeout 2400 -- 'timeout 2350 bash tools/gpu_suite.sh r02'
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed" gpurun_out/${TAG}_pytest.log | tail -2
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
echo "bench rc=$?"
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
echo "bench ref rc=$?"
NOX="--no-cpu-baseline --no-objf-check --no-ctc-roofline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 $NOX > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 250 ncu --set full --clock-control none --import-source on -k regex:rec_tc_fwd -s 10 -c 1 -o gpurun_out/${TAG}_prof_rec -f \
    python bench.py --steps 2 --warmup 1 $NOX > gpurun_out/${TAG}_ncu_rec.log 2>&1; echo "ncu rec fwd rc=$?"
timeout 250 ncu --set full --clock-control none --import-source on -k regex:rec_tc_bwd -s 5 -c 1 -o gpurun_out/${TAG}_prof_recbwd -f \
    python bench.py --steps 2 --warmup 1 $NOX > gpurun_out/${TAG}_ncu_recbwd.log 2>&1; echo "ncu rec bwd rc=$?"
timeout 100 python tools/gemm_pair_time.py > gpurun_out/${TAG}_gemm_pair_time.log 2>&1; echo "gemm time rc=$?"
timeout 250 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_pair -s 5 -c 1 -o gpurun_out/${TAG}_prof_gemm -f \
    python tools/gemm_pair_time.py > gpurun_out/${TAG}_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 60 python tools/ctc_stress_time.py 32 2 > gpurun_out/${TAG}_ctc32.log 2>&1; echo "ctc32 rc=$?"
timeout 250 ncu --set full --clock-control none --import-source on -k regex:ctc_ -s 6 -c 6 -o gpurun_out/${TAG}_prof_ctc -f \
    python tools/ctc_stress_time.py 32 2 > gpurun_out/${TAG}_ncu_ctc.log 2>&1; echo "ncu ctc rc=$?"
B200CTC_PROFILE=1 timeout 100 python tools/ctc_roofline.py 256 2>&1 | grep b200ctc | tail -1 > gpurun_out/${TAG}_ctc_per_kernel.log
timeout 100 python tools/ctc_time.py 1 4 > gpurun_out/${TAG}_ctc_small.log 2>&1
python - "$TAG" <<'PY'
import json,sys
tag=sys.argv[1]
try:
    d=json.loads([l for l in open("gpurun_out/%s_bench_n1.json"%tag) if l.startswith("{")][0])
    print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],"objf_check",d["objf_check"]["rel_diff"],"ctc",d["ctc_roofline"]["frac"],d["ctc_roofline"]["ms_per_call"],"gemm",d["gemm_roofline"]["frac"],d["gemm_roofline"]["achieved"],"fs3",d["fs3"]["ms_per_step"],"gru",d["configs3_gru"]["ms_per_step"],"cpu",d["cpu_baseline"]["value"])
except Exception as e: print("ERR",e)
PY
ls -la gpurun_out/${TAG}_*.ncu-rep 2>/dev/null
