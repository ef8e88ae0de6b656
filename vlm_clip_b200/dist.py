"""Data-parallel plumbing of the contrastive step (new relative to the reference, which is single-process;
SURVEY.md §8e).  One process per GPU, torch.distributed (NCCL over NVLink on the box, gloo in the CPU tests).

  gather_features   all-gather of the UN-normalised [B_loc, P] text and image features; returns the global matrices
                    and this rank's first row.  Every rank then evaluates its two STRIPS of the N x N logit matrix
                    (its text rows x all images, its image rows x all texts: ops.clip_loss(exchange=group)), the ranks
                    exchange the 2 B_loc log-sum-exps these give (one more small all-gather), and each rank
                    differentiates only its own rows (both the row- and the column-softmax terms), which is exactly
                    dL_global / d(local features) — no reduce-scatter.
  allreduce_sum_    adapter gradients: SUM, not mean (the 1/N already lives in the loss); clip AFTER this.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def gather_features(text_features: torch.Tensor, image_features: torch.Tensor, group=None):
    """All-gather of the un-normalised local features into the global [N, P] matrices (two collectives writing straight
    into their outputs: no pack / unpack copies around them).  Returns (text_all, image_all, first global row)."""
    ws, rank = world(group)
    if ws == 1:
        return text_features.detach(), image_features.detach(), 0
    B, P = text_features.shape
    outs = []
    for f in (text_features, image_features):
        f = f.detach()
        if not f.is_contiguous():
            f = f.contiguous()
        g = torch.empty((ws * B, P), device=f.device, dtype=f.dtype)
        dist.all_gather_into_tensor(g, f, group=group)
        outs.append(g)
    return outs[0], outs[1], rank * B


def allreduce_sum_(flat_grad: torch.Tensor, group=None):
    ws, _ = world(group)
    if ws > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad
