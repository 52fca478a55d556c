timeout 700 python -m pytest tests/test_rnn_tc_gpu.py tests/test_bench_size_gpu.py tests/test_train_step_gpu.py tests/test_rnn_gpu.py tests/test_cudnn_compat_gpu.py -m gpu -q --timeout 300 2>&1 | tail -4
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ctc-roofline --no-objf-check > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/t_bench.json") if l.startswith("{")][0])
print("ms",d["ms_per_step"],"e2e",d["e2e"]["value"],d["roofline"].get("ms_per_step_by_kernel"), d["configs3_gru"]["ms_per_step"])
PY
