timeout 300 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q --timeout 60 2>&1 | grep -E "Error|assert |error|max err|passed|failed" | head -12
