#!/bin/bash
# tuning aid: per-kernel times (B200CTC_PROFILE=1, serialised) and the concurrent timeline (=2) of the
# CTC call on the configs[4] workload.  Usage: bash tools/ctc_sweep.sh [B]
B=${1:-256}
B200CTC_PROFILE=1 python tools/ctc_roofline.py $B 2>&1 | grep -E "b200ctc" | tail -1
for g in 1 2 4; do
  B200CTC_GROUPS=$g B200CTC_PROFILE=2 python tools/ctc_roofline.py $B 2>&1 | grep -E "b200ctc|ms_per_call" | tail -$((g+1)) | cut -c1-330
done
