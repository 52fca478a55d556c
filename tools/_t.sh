timeout 600 python -m pytest tests/test_ctc_gpu.py -m gpu -q --timeout 300 2>&1 | tail -3
B200CTC_PROFILE=1 python tools/ctc_roofline.py 256 2>&1 | grep -E "b200ctc" | tail -1 | cut -c1-300
for g in 1 2 3 4; do echo "groups $g: $(B200CTC_GROUPS=$g python tools/ctc_roofline.py 256 2>&1 | grep -E "ms_per_call" | tail -1 | sed -e 's/.*"frac": \([0-9.]*\).*"ms_per_call": \([0-9.]*\).*within_tolerance": \([a-z]*\).*/frac \1 ms \2 ok \3/')"; done
python tools/ctc_time.py 1 4 2>&1 | cut -c1-120
