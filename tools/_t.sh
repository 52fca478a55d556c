timeout 600 python -m pytest tests/test_ctc_gpu.py -m gpu -q --timeout 300 2>&1 | tail -5
B200CTC_PROFILE=1 python tools/ctc_roofline.py 256 2>&1 | grep -E "b200ctc" | tail -1 | cut -c1-300
python tools/ctc_roofline.py 256 2>&1 | grep -E "ms_per_call" | tail -1 | sed -e 's/.*"frac": \([0-9.]*\).*"ms_per_call": \([0-9.]*\).*"parity_fp64_oracle_full_size": \(.*\), "l2.*/frac \1 ms \2 parity \3/'
B200CTC_NA=4 B200CTC_PROFILE=1 python tools/ctc_roofline.py 256 2>&1 | grep -E "b200ctc" | tail -1 | cut -c1-300
