"""Data-parallel plumbing of the contrastive step (new relative to the reference, which is single-process;
SURVEY.md §8e).  One process per GPU, torch.distributed (NCCL over NVLink on the box, gloo in the CPU tests).

  gather_features   all-gather of the UN-normalised [B_loc, P] text and image features as one [B_loc, 2P] message;
                    returns the global matrices and this rank's first row.  Every rank then evaluates the full
                    N x N loss and differentiates only its own rows (both the row- and the column-softmax terms),
                    which is exactly dL_global / d(local features) — no reduce-scatter.
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
    ws, rank = world(group)
    if ws == 1:
        return text_features.detach(), image_features.detach(), 0
    B, P = text_features.shape
    both = torch.cat([text_features.detach(), image_features.detach()], dim=1).contiguous()
    gathered = torch.empty((ws * B, 2 * P), device=both.device, dtype=both.dtype)
    dist.all_gather_into_tensor(gathered, both, group=group)
    return gathered[:, :P].contiguous(), gathered[:, P:].contiguous(), rank * B


def allreduce_sum_(flat_grad: torch.Tensor, group=None):
    ws, _ = world(group)
    if ws > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad
