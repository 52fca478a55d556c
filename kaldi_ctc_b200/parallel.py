"""Utterance-sharded data parallelism: host-side plumbing only.

The reference has no collective backend: it runs one nnet2-ctc-train-simple per GPU on its own egs
archive and averages the models on disk (egs/wsj/s5/steps/ctc/train.sh:408-435).  Here every rank
takes its own utterances (the local slab keeps the t*B_local + b layout), weight gradients are
SUMMED over ranks before the element-wise clip and the update, which makes N ranks x B utterances
equal to one minibatch of N*B utterances.
"""
import torch.distributed as dist


def shard_utterances(num_utts, world, rank):
    """Indices of the utterances rank `rank` trains on: contiguous, sizes differing by at most one."""
    base, rem = divmod(num_utts, world)
    lo = rank * base + min(rank, rem)
    return list(range(lo, lo + base + (1 if rank < rem else 0)))


class GradientReducer:
    """Issues one asynchronous all-reduce(sum) per gradient tensor as soon as it exists and applies
    the caller's update when the reduction has landed (top layer first, so the transfers overlap the
    lower layers' backward)."""

    def __init__(self, group=None):
        self.group = group
        self.pending = []

    def submit(self, tensors, on_done):
        works = [dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True) for t in tensors]
        self.pending.append((works, on_done))

    def submit_max(self, flag):
        """all-reduce(max) of a small flag tensor, ahead of the gradients on the same communicator."""
        self.pending.append(([dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group, async_op=True)],
                             lambda: None))

    def finish(self):
        for works, on_done in self.pending:
            for w in works:
                w.wait()
            on_done()
        self.pending = []


def reduce_scalar_sum(value, device="cpu"):
    """Sum of a python float over ranks (objective / frame counts for logging)."""
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
