"""Host->device staging for the train loop (the only boundary crossing of the hot path: trainer.py:75-78).

`DevicePrefetcher` wraps any iterable of batch dicts (a DataLoader) and copies batch k+1 from pinned host
memory on a side stream while step k computes, so the H2D copy of a 154 MB fp32 pixel batch (B=256) hides
behind the ~10 ms step instead of preceding it.  `SyntheticPairs` produces the SURVEY.md §8d synthetic inputs.
"""
from __future__ import annotations

import torch


class DevicePrefetcher:
    def __init__(self, loader, device, depth: int = 2):
        self.loader = loader
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.depth = depth

    def __len__(self):
        return len(self.loader)

    def _stage(self, batch):
        out = {}
        with torch.cuda.stream(self.stream):
            for k, v in batch.items():
                if isinstance(v, torch.Tensor):
                    if not v.is_cuda and not v.is_pinned():
                        v = v.pin_memory()
                    out[k] = v.to(self.device, non_blocking=True)
                else:
                    out[k] = v
        ev = torch.cuda.Event()
        ev.record(self.stream)
        return out, ev

    def __iter__(self):
        it = iter(self.loader)
        queue = []
        try:
            for _ in range(self.depth):
                queue.append(self._stage(next(it)))
        except StopIteration:
            pass
        while queue:
            batch, ev = queue.pop(0)
            torch.cuda.current_stream(self.device).wait_event(ev)
            for v in batch.values():
                if isinstance(v, torch.Tensor):
                    v.record_stream(torch.cuda.current_stream(self.device))
            try:
                queue.append(self._stage(next(it)))
            except StopIteration:
                pass
            batch["inputs_ready"] = ev  # lets the frozen towers start without waiting for the consumer's stream
            yield batch


class SyntheticPairs(torch.utils.data.Dataset):
    """Image/caption pairs of SURVEY.md §8d: pixels ~ N(0,1) fp32; ids ~ U{3..49405}, BOS first, EOS last; mask 1."""

    def __init__(self, n: int, seed: int = 2, image: int = 224, seq: int = 77):
        g = torch.Generator().manual_seed(seed)
        self.pix = torch.randn(n, 3, image, image, generator=g)
        self.ids = torch.randint(3, 49406, (n, seq), generator=g)
        self.ids[:, 0] = 49406
        self.ids[:, -1] = 49407
        self.mask = torch.ones(n, seq, dtype=torch.int64)

    def __len__(self):
        return self.pix.shape[0]

    def __getitem__(self, i):
        return {"input_ids": self.ids[i], "attention_mask": self.mask[i], "pixel_values": self.pix[i]}
