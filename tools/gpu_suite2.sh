#!/bin/bash
# new tests of this round + ncu captures of the backward recurrent kernel and the CTC kernels
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_clip_gradient_gpu.py tests/test_reference_linked_gpu.py tests/test_train_step_gpu.py tests/test_cudnn_compat_gpu.py -q --timeout 200 > gpurun_out/r02b_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/r02b_pytest.log | tail -6
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ctc-roofline --no-objf-check > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"
timeout 250 ncu --set full --clock-control none --import-source on -k regex:rec_tc_bwd -s 5 -c 1 -o gpurun_out/r02_prof_recbwd -f \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-objf-check --no-ctc-roofline > gpurun_out/r02b_ncu_recbwd.log 2>&1; echo "ncu bwd rc=$?"
timeout 60 python tools/ctc_stress_time.py 32 2 > gpurun_out/r02b_ctc32.log 2>&1
timeout 250 ncu --set full --clock-control none -k regex:ctc_ -s 6 -c 3 -o gpurun_out/r02_prof_ctc -f \
    python tools/ctc_stress_time.py 32 2 > gpurun_out/r02b_ncu_ctc.log 2>&1; echo "ncu ctc rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02b_bench.json") if l.startswith("{")][0])
print(d["ms_per_step"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["ms_per_step_by_kernel"])
PY
