#!/bin/bash
mkdir -p gpurun_out
( B200CTC_STREAM=1 timeout 900 python -m pytest tests/test_ctc_gpu.py -q --timeout 600 -x 2>&1 | tail -15 ) > gpurun_out/c3_pytest_stream.log 2>&1
( timeout 900 python -m pytest tests/test_bench_size_gpu.py -q --timeout 600 -k "spot_check" 2>&1 | tail -15 ) >> gpurun_out/c3_pytest_stream.log 2>&1
tail -12 gpurun_out/c3_pytest_stream.log
for B in 256 128 64; do
  B200CTC_STREAM=1 B200CTC_PROFILE=1 timeout 300 python tools/ctc_stress_time.py $B 3
done > gpurun_out/c3_ctc_prof.log 2>&1
for NA in 3 4; do B200CTC_STREAM=1 B200CTC_STREAM_NA=$NA B200CTC_PROFILE=1 timeout 300 python tools/ctc_stress_time.py 256 3; done > gpurun_out/c3_ctc_na.log 2>&1
timeout 300 python tools/ctc_stress_time.py 256 5 > gpurun_out/c3_ctc_default.log 2>&1
grep -h "streaming\|alg_GBps" gpurun_out/c3_ctc_prof.log | cut -c1-200 | uniq -w 60
grep -h "streaming\|alg_GBps" gpurun_out/c3_ctc_na.log | cut -c1-200 | uniq -w 60
cat gpurun_out/c3_ctc_default.log | cut -c1-200
