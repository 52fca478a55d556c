#!/bin/bash
mkdir -p gpurun_out
B200CTC_STREAM=1 timeout 60 python tools/ctc_stress_time.py 64 2 > gpurun_out/c6_plain.log 2>&1 || exit 1
B200CTC_STREAM=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:ctc_stream -s 4 -c 2 \
   -o gpurun_out/prof_ctc_stream_r02 -f python tools/ctc_stress_time.py 64 2 > gpurun_out/c6_ncu.log 2>&1
tail -3 gpurun_out/c6_ncu.log; ls -la gpurun_out/*.ncu-rep
