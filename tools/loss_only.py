"""Runs the three launches of the strip-wise contrastive loss (csrc/clip_loss.cu) at BASELINE config 3's size
(N = 4096 global pairs, P = 768, this rank's 512 rows): the target of `ncu -k regex:clip_` captures.  Development tool."""
import sys

import torch

sys.path.insert(0, ".")
from vlm_clip_b200 import _native as N  # noqa: E402

lib = N.load()
dev = torch.device("cuda:0")
Nn, P, nl = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (4096, 768, 512)
g = torch.Generator(device="cuda").manual_seed(1)
t = torch.randn(Nn, P, device=dev, generator=g)
i = torch.randn(Nn, P, device=dev, generator=g)
tn, im = torch.empty_like(t), torch.empty_like(i)
blk = torch.zeros(Nn // nl, 2 * nl + 1, device=dev)
counters = torch.zeros(int(lib.vlmclip_clip_loss_counters(nl)), device=dev, dtype=torch.int32)
state = torch.empty(int(lib.vlmclip_clip_loss_state_size(Nn, P, nl)), device=dev)
ws = torch.empty(int(lib.vlmclip_clip_loss_bwd_workspace(Nn, P, nl)), device=dev)
dt, di = torch.empty(nl, P, device=dev), torch.empty(nl, P, device=dev)
for r in range(Nn // nl):  # every "rank's" LSE block, so that the backward sees a complete exchange
    N.check(lib.vlmclip_clip_loss_fwd(N.ptr(t), N.ptr(i), 100.0, N.ptr(tn), N.ptr(im), None, N.ptr(blk[r]), N.ptr(blk[r][2 * nl:]),
                                      N.ptr(state), N.ptr(counters), Nn, P, r * nl, nl, N.stream()), "fwd")
for _ in range(3):
    N.check(lib.vlmclip_clip_loss_fwd(N.ptr(t), N.ptr(i), 100.0, N.ptr(tn), N.ptr(im), None, N.ptr(blk[0]), N.ptr(blk[0][2 * nl:]),
                                      N.ptr(state), N.ptr(counters), Nn, P, 0, nl, N.stream()), "fwd")
    N.check(lib.vlmclip_clip_loss_bwd(N.ptr(tn), N.ptr(im), N.ptr(blk), 2 * nl + 1, nl, 100.0, N.ptr(dt), N.ptr(di), None,
                                      N.ptr(state), N.ptr(counters), N.ptr(ws), Nn, P, 0, nl, 0, nl, N.stream()), "bwd")
torch.cuda.synchronize()
print("loss", float(blk[:, 2 * nl].sum()), "finite grads", bool(torch.isfinite(dt).all() and torch.isfinite(di).all()))
