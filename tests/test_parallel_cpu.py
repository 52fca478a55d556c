"""Host-side data-parallel logic on CPU with gloo, world_size 2 (no GPU): utterance sharding and
'sum over ranks, then clip, then update' == one big minibatch.  The per-rank gradients come from
the oracle (test infrastructure); the reduction goes through kaldi_ctc_b200.parallel."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kaldi_ctc_b200 import parallel


def test_shard_utterances_partitions_the_minibatch():
    for n, w in [(16, 2), (16, 8), (10, 4), (3, 4), (48, 3)]:
        parts = [parallel.shard_utterances(n, w, r) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle
    rng = np.random.default_rng(0)                 # identical data on every rank
    D, H, B, T = 6, 8, 6, 9
    w = (rng.standard_normal(pyoracle.rnn_param_count(2, True, 1, D, H)) * 0.4).astype(np.float32)
    x = rng.standard_normal((T, B, D)).astype(np.float32)
    dy = rng.standard_normal((T, B, 2 * H)).astype(np.float32) * 4
    mine = parallel.shard_utterances(B, world, rank)
    xs = np.ascontiguousarray(x[:, mine]).reshape(T * len(mine), D)
    dys = np.ascontiguousarray(dy[:, mine]).reshape(T * len(mine), 2 * H)
    _, _, dw = pyoracle.rnn(2, True, 1, H, xs, w, len(mine), dy=dys, dtype=np.float64)
    g = torch.from_numpy(dw.copy())
    applied = {}
    red = parallel.GradientReducer()
    red.submit([g], lambda: applied.setdefault("w", w + 0.1 * np.clip(g.numpy(), -5, 5)))
    red.finish()
    tot = parallel.reduce_scalar_sum(float(len(mine)))
    if rank == 0:
        _, _, dw_full = pyoracle.rnn(2, True, 1, H, x.reshape(T * B, D), w, B, dy=dy.reshape(T * B, 2 * H),
                                     dtype=np.float64)
        out["err"] = float(np.abs(g.numpy() - dw_full).max())
        out["clipped"] = bool((np.abs(dw_full) > 5).any())
        out["upd_err"] = float(np.abs(applied["w"] - (w + 0.1 * np.clip(dw_full, -5, 5))).max())
        out["tot"] = tot
    dist.destroy_process_group()


def test_sum_then_clip_equals_one_big_minibatch_gloo_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out["tot"] == 6.0
    assert out["err"] < 1e-10          # sum of the per-rank gradients == gradient of the whole minibatch
    assert out["clipped"]              # the clamp is exercised, so clip-after-sum matters
    assert out["upd_err"] < 1e-10
