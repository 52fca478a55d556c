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
    lower layers' backward).  Updates lag their reduction by ONE submission: when the next component's
    gradient is submitted (a recurrent layer later, i.e. long after the previous transfer has finished),
    the previous components' reductions are waited for and their updates applied -- only the last
    component's reduction and update are left for finish(), instead of every update of the step."""

    def __init__(self, group=None):
        self.group = group
        self.pending = []

    def _drain(self, entries):
        for works, on_done in entries:
            for w in works:
                w.wait()
            on_done()

    def submit(self, tensors, on_done, drain=True):
        """drain=False: only queue (for a caller that is on a stream which must not wait for transfers)."""
        works = [dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True) for t in tensors]
        if not drain:
            self.pending.append((works, on_done))
            return
        older, self.pending = self.pending, [(works, on_done)]
        self._drain(older)   # (after the new transfer has been issued: it starts while the older updates run)

    def submit_max(self, flag):
        """all-reduce(max) of a small flag tensor, ahead of the gradients on the same communicator."""
        self.pending.append(([dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group, async_op=True)],
                             lambda: None))

    def finish(self):
        older, self.pending = self.pending, []
        self._drain(older)


def reduce_scalar_sum(value, device="cpu"):
    """Sum of a python float over ranks (objective / frame counts for logging)."""
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
