"""CFG-parallel x Ulysses: the conditional and unconditional forwards of one guidance step (wan/text2video.py:509-562) run
on two halves of the world, each half a Ulysses sequence-parallel group of world/2 ranks.

Why: Wan2.1-1.3B has 12 heads, which 8 GPUs cannot split (SURVEY §7 "12 ∤ 8", §8e "CFG cond/uncond can additionally be split
over two GPU groups"); two groups of 4 can, and every exchange then spans half as many peers.  The only collective this adds is
one 2-rank all-gather of the fp32 prediction per step (8.4 MB at 832x480x81) between rank r and rank r + world/2 — a real
exchange step of the path: both halves need cond AND uncond for the CFG(-Zero*) combine and the scheduler update, which stay
replicated.  The reference's own multi-GPU path (xdit_context_parallel.py) only has the Ulysses half; the split of the
guidance batch is xDiT's `cfg_parallel`, which the reference's launcher does not expose.
"""
from __future__ import annotations

from typing import Tuple

import torch


class CfgParallel:
    def __init__(self, world_group=None):
        import torch.distributed as dist
        self.dist = dist
        world = dist.get_world_size(world_group)
        rank = dist.get_rank(world_group)
        if world % 2:
            raise ValueError("CfgParallel needs an even number of ranks")
        half = world // 2
        self.world, self.rank, self.half = world, rank, half
        self.branch = rank // half                     # 0: conditional forward, 1: unconditional forward
        # every rank has to take part in the creation of every group
        sp = [dist.new_group(list(range(b * half, (b + 1) * half))) for b in range(2)]
        pairs = [dist.new_group([r, r + half]) for r in range(half)]
        self.sp_group = sp[self.branch]                # hand this to WanModel(sp_group=...)
        self.pair_group = pairs[rank % half]

    def select(self, cond, uncond):
        """the input (context, ...) of this rank's branch"""
        return cond if self.branch == 0 else uncond

    def exchange(self, pred: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """this rank's prediction -> (cond prediction, uncond prediction) on every rank"""
        bufs = [torch.empty_like(pred), torch.empty_like(pred)]
        self.dist.all_gather(bufs, pred.contiguous(), group=self.pair_group)   # pair-group rank 0 is the cond half
        return bufs[0], bufs[1]
