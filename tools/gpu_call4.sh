#!/bin/bash
# every step under its own short timeout: a hung kernel costs seconds, not the whole call
mkdir -p gpurun_out
( B200CTC_STREAM=1 timeout 180 python -m pytest tests/test_ctc_gpu.py -q --timeout 60 -x 2>&1 | tail -8 ) > gpurun_out/c4_pytest_stream.log 2>&1
tail -4 gpurun_out/c4_pytest_stream.log
for B in 256 64; do
  B200CTC_LIB=tools/build/libb200ctc_prof.so B200CTC_STREAM=1 B200CTC_PROFILE=1 timeout 100 python tools/ctc_stress_time.py $B 2
done > gpurun_out/c4_ctc_pc.log 2>&1
grep -h -A6 "streaming" gpurun_out/c4_ctc_pc.log | tail -32
timeout 100 python tools/ctc_stress_time.py 256 5 > gpurun_out/c4_ctc_default.log 2>&1; cut -c1-200 gpurun_out/c4_ctc_default.log
