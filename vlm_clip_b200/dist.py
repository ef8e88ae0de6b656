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


class BucketedGradAllReduce:
    """SUM all-reduce of a flat gradient arena in buckets that start while the backward pass is still running
    (BASELINE config 5: 151 M parameters = 604 MB of fp32 gradients per step; SURVEY.md 8e "bucket + overlap").

    The full-fine-tune backward is hand-written and walks the layers from the last to the first
    (finetune._encoder_bwd); the 16 parameters of an encoder layer are contiguous in the optimiser's arena (they are
    registered in `named_parameters()` order).  `sink(params, grads)` is called by the backward as soon as a layer's
    gradients exist: it writes them into the arena views (`param.grad`) and starts an asynchronous NCCL all-reduce of
    the layer's arena range, which then runs under the reverse pass of the layers below.  `finish()` reduces what no
    bucket covered (embeddings, projections, adapters, logit_scale: whatever autograd accumulated on its own) and makes
    the caller's stream wait for every bucket."""

    def __init__(self, optimizer, group=None):
        self.opt = optimizer
        self.group = group
        self.offset = {id(p): (off, p.numel()) for p, off in zip(optimizer.params, optimizer.offsets)}
        self.reset()

    def reset(self):
        self.works = []
        self.covered = []  # arena ranges [start, end) already handed to NCCL

    def sink(self, params, grads):
        lo, hi = None, None
        for p, g in zip(params, grads):
            off, n = self.offset[id(p)]
            p.grad.copy_(g.view_as(p.grad))  # the arena was zeroed by zero_grad(): this IS the accumulated gradient
            lo = off if lo is None else min(lo, off)
            hi = off + n if hi is None else max(hi, off + n)
        hi = (hi + 3) // 4 * 4  # parameters start on 16-byte boundaries; the padding is zero on every rank
        self.covered.append((lo, hi))
        self.works.append(dist.all_reduce(self.opt.grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        pos = 0
        for lo, hi in sorted(self.covered) + [(self.opt.n, self.opt.n)]:
            if lo > pos:
                self.works.append(dist.all_reduce(self.opt.grad[pos:lo], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            pos = max(pos, hi)
        for w in self.works:
            w.wait()  # the current stream waits for the collective's stream; the host does not block
        n = len(self.works)
        self.reset()
        return n
