"""CTC roofline measurement of bench.py on its own (configs[4], device-generated activations)."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
pk, src = bench.peaks()
print(json.dumps(bench.ctc_roofline("cuda:0", pk, src, B=int(sys.argv[1]) if len(sys.argv) > 1 else 256)))
