"""Data-parallel parity on real GPUs (run under torchrun, one rank per GPU):
N ranks x B utterances with summed gradients must give the weights of ONE process training on the
N*B utterances (sum-then-clip).  Equal-length utterances, so that padding is identical.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py [fp32|tensor]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from kaldi_ctc_b200 import nnet, parallel, rnn, synth  # noqa: E402

math = rnn.MATH_TENSOR if (len(sys.argv) > 1 and sys.argv[1] == "tensor") else rnn.MATH_FP32
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = "cuda:%d" % int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(dev)
dist.init_process_group("nccl")
Bl, T = 4, 48
spec = synth.ModelSpec(mode=2, layers=3, D=16, H=64, A=20, learning_rate=0.01, param_stddev=0.2)
blobs, aw, ab = synth.model_weights(spec, 3)
x, fl, L, Tl = synth.features(Bl * world, spec.D, T, T, 3, 6, spec.A, seed=11)   # [T*Btot, D], row t*Btot+b
Bt = Bl * world
mine = parallel.shard_utterances(Bt, world, rank)
x3 = x.reshape(T, Bt, spec.D)
offs = np.concatenate([[0], np.cumsum(L)])
xl = np.ascontiguousarray(x3[:, mine]).reshape(T * Bl, spec.D)
fll = np.concatenate([fl[offs[b]:offs[b + 1]] for b in mine])
up = nnet.NnetCtcUpdater(spec, blobs, aw, ab, Bl, T, device=dev, math=math, world=world)
steps = 2
for _ in range(steps):
    objf = up.ComputeForMinibatch(torch.from_numpy(xl).pin_memory(), T, fll, L[mine], Tl[mine])
tot = parallel.reduce_scalar_sum(objf, device="cuda")
ok = True
if rank == 0:
    ref = nnet.NnetCtcUpdater(spec, blobs, aw, ab, Bt, T, device=dev, math=math, world=1)
    for _ in range(steps):
        objf_ref = ref.ComputeForMinibatch(torch.from_numpy(x).pin_memory(), T, fl, L, Tl)
    tol = 2e-5 if math == rnn.MATH_FP32 else 2e-3
    worst = 0.0
    for a, b in zip(up.rnns, ref.rnns):
        worst = max(worst, float((a.filter_params_ - b.filter_params_).abs().max()))
    worst = max(worst, float((up.affine.linear_params_ - ref.affine.linear_params_).abs().max()))
    ok = worst < tol and abs(tot - objf_ref) < 1e-4 * abs(objf_ref) * (1 if math == rnn.MATH_FP32 else 50)
    print("DP_CHECK world=%d math=%s max|w_dp - w_single|=%.3g (tol %.1g) objf %.4f vs %.4f -> %s" %
          (world, "fp32" if math == rnn.MATH_FP32 else "tensor", worst, tol, tot, objf_ref, "OK" if ok else "FAIL"))
# every rank must hold identical weights
w = torch.cat([c.filter_params_ for c in up.rnns])
lo, hi = w.clone(), w.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
same = bool((lo == hi).all())
if rank == 0:
    print("DP_CHECK replicas bit-identical:", same)
dist.destroy_process_group()
sys.exit(0 if (ok and same) else 1)
