"""Guidance-condition parallelism for the LTX loop ("the 3 guidance conds are independent forwards").

One denoise step of `LTXVideoPipeline.__call__` (pipeline_ltx_video.py:1103-1256) runs the transformer on a batch of num_conds
copies of the latents — [unconditional, text, STG-perturbed] — that only differ in the prompt rows and the skip-layer mask, and only
meet again in the guidance arithmetic.  Here the condition rows are split over the ranks of a process group: every rank runs the
transformer on its own rows, the predictions are exchanged (one broadcast per owning rank: 1.5 MB of bf16 per condition and video at
768x512x121), and the guidance + scheduler kernel runs replicated, so every rank holds the same latents bit for bit after every step.
That turns N GPUs into LATENCY for ONE video (3 conditions: 1/3 of the step on 3 GPUs, 2/3 on 2), where replicas only add throughput.
The exchange is the path's real exchange step, so it is a collective (NCCL broadcast); nothing else crosses ranks.
"""
from __future__ import annotations

from typing import List, Tuple


def cond_partition(num_conds: int, ranks: int) -> List[Tuple[int, int]]:
    """Contiguous [lo, hi) condition ranges per rank: the first ranks take ceil(num_conds / ranks) conditions, ranks beyond the
    number of conditions take none (they still follow the latents)."""
    per = -(-num_conds // max(ranks, 1))
    out = []
    for r in range(ranks):
        lo = min(r * per, num_conds)
        out.append((lo, min(lo + per, num_conds)))
    return out


class CondParallel:
    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._global = [dist.get_global_rank(group, r) if group is not None else r for r in range(self.size)]

    def ranges(self, num_conds: int):
        return cond_partition(num_conds, self.size)

    def exchange(self, pred_full, bsz: int, ranges) -> None:
        """pred_full [num_conds * bsz, N, C]: every owner's rows -> every rank (in place)."""
        for r, (lo, hi) in enumerate(ranges):
            if hi > lo:
                self.dist.broadcast(pred_full[lo * bsz:hi * bsz], src=self._global[r], group=self.group)

    def any_flag(self, flag: bool, device) -> bool:
        """OR of a host flag over the group (the `_interrupt` poll: every rank must leave the loop at the same step)."""
        import torch
        t = torch.tensor([1 if flag else 0], device=device, dtype=torch.int32)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return bool(int(t.item()))
