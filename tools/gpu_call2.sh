#!/bin/bash
mkdir -p gpurun_out
( B200CTC_STREAM=1 timeout 900 python -m pytest tests/test_ctc_gpu.py -q --timeout 600 -x 2>&1 | tail -15 ) > gpurun_out/c2_pytest_stream.log 2>&1
( timeout 900 python -m pytest tests/test_bench_size_gpu.py -q --timeout 600 -k "spot_check" 2>&1 | tail -15 ) >> gpurun_out/c2_pytest_stream.log 2>&1
( timeout 600 python -m pytest tests/test_cudnn9_crosscheck_gpu.py -q --timeout 600 2>&1 | tail -5 ) >> gpurun_out/c2_pytest_stream.log 2>&1
tail -30 gpurun_out/c2_pytest_stream.log
for B in 256 128 64 32; do
  B200CTC_STREAM=1 B200CTC_PROFILE=1 timeout 300 python tools/ctc_stress_time.py $B 3
  B200CTC_STREAM=0 B200CTC_PROFILE=1 timeout 300 python tools/ctc_stress_time.py $B 3
done > gpurun_out/c2_ctc_prof.log 2>&1
for NA in 3 5; do B200CTC_STREAM=1 B200CTC_STREAM_NA=$NA timeout 300 python tools/ctc_stress_time.py 256 5; done > gpurun_out/c2_ctc_na.log 2>&1
timeout 300 python tools/ctc_stress_time.py 256 5 > gpurun_out/c2_ctc_default.log 2>&1
grep -h "streaming\|rowstats\|alg_GBps" gpurun_out/c2_ctc_prof.log | cut -c1-260
cat gpurun_out/c2_ctc_na.log gpurun_out/c2_ctc_default.log | cut -c1-200
