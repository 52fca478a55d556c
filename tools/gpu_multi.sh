#!/bin/bash
# bench.py under torchrun on N GPUs of one box (both arms, as the driver launches them).  Usage: tools/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
echo "bench n=$N rc=$?"
python - $N <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open("gpurun_out/r02_bench_n%s.json"%n) if l.startswith("{")][0])
    print("n",d["n_gpus"],"value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],"tail",d["side_stream_wait_ms_per_step"])
    print("fs3",d["fs3"]); print("strong",d["strong"])
except Exception as e:
    print("ERR",e); print(open("gpurun_out/r02_bench_n%s.err"%n).read()[-1500:])
PY
