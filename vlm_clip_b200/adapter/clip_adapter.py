"""B200-native mirrors of the reference's sequence adapters (reference: adapter/clip_adapter.py).

Same class names, constructor signatures, attribute names and state-dict keys (`down_project`, `up_project`,
`layer_norm`; fixture test_checkpoints/test_adapter.pt), so checkpoints are interchangeable and
`trainer.py:40-43`'s `"adapter" in name` filter keeps working.  The arithmetic

    y = LayerNorm(up_project(GELU(down_project(x))) + x)          (adapter/clip_adapter.py:17-23, 144-150)

runs in ONE fused CUDA kernel (vlmclip_adapter_fwd) with a matching backward that produces the adapter's own
gradients (vlmclip_adapter_bwd).  Parameters stay fp32 nn.Parameters owned by torch.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _native as N
from .. import ops


class _SeqBottleneckAdapter(nn.Module):
    def __init__(self, hidden_size, adapter_size):
        super().__init__()
        self.down_project = nn.Linear(hidden_size, adapter_size)
        self.activation = nn.GELU()
        self.up_project = nn.Linear(adapter_size, hidden_size)
        self.layer_norm = nn.LayerNorm(hidden_size)

    def _run(self, x2d, rows=None, ldx=None):
        ln = self.layer_norm
        return ops.adapter(x2d, self.down_project.weight, self.down_project.bias, self.up_project.weight,
                           self.up_project.bias, ln.weight, ln.bias, act=N.ACT_GELU_ERF, post=N.POST_RESIDUAL_LN,
                           eps=ln.eps, rows=rows, ldx=ldx)

    def forward(self, hidden_states):
        """hidden_states: [..., hidden_size] (fp32, or the towers' bf16) on CUDA -> fp32, same shape."""
        D = self.down_project.in_features
        if hidden_states.shape[-1] != D:
            raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied (last dim {hidden_states.shape[-1]} != {D})")
        x = hidden_states
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        x2d = x.reshape(-1, D)
        if not x2d.is_contiguous():
            x2d = x2d.contiguous()
        y = self._run(x2d)
        return y.view(*hidden_states.shape[:-1], D)

    def forward_token0(self, hidden_flat, batch: int, seq: int):
        """Result-identical fast path of `self(h)[:, 0, :]` (model_m.py:102,122): the adapter is position-wise, so
        only row 0 of every sequence is evaluated.  hidden_flat: contiguous [batch*seq, hidden] (bf16 or fp32)."""
        D = self.down_project.in_features
        return self._run(hidden_flat, rows=batch, ldx=seq * D)


class TextAdapter(_SeqBottleneckAdapter):
    """Adapter module for the CLIP text encoder (reference: adapter/clip_adapter.py:4-23)."""


class VisionAdapter(_SeqBottleneckAdapter):
    """Adapter module for the CLIP vision encoder (reference: adapter/clip_adapter.py:131-150)."""


class SharedMHSAttentionAdapter(nn.Module):
    """Cross-modal adapter (reference: adapter/clip_adapter.py:69-128).

    Parameter layout and state-dict keys match the reference (text_proj, image_proj, cross_attn, norm1..3, mlp).
    The reference can only execute this module at batch size 1 (model_m.py:96-100 passes a batch-1 K/V to
    nn.MultiheadAttention) and it hard-codes 512/768 widths.  forward() runs the bf16 tensor-core
    inference path under no_grad in eval mode and the fp32 trainable path (fused autograd node, hand-written backward,
    adapter/shared_train.py) otherwise; both evaluate every text row on its own against the shared table.
    """

    def __init__(self, text_input_size=512, image_input_size=768, hidden_size=512, num_heads=8, dropout=0.1):
        super().__init__()
        self.text_proj = nn.Linear(text_input_size, hidden_size)
        self.image_proj = nn.Linear(image_input_size, hidden_size)
        self.cross_attn = nn.MultiheadAttention(embed_dim=hidden_size, num_heads=num_heads, dropout=dropout,
                                                batch_first=True)
        self.norm1 = nn.LayerNorm(hidden_size)
        self.norm2 = nn.LayerNorm(hidden_size)
        self.norm3 = nn.LayerNorm(hidden_size)
        self.mlp = nn.Sequential(nn.Linear(hidden_size, hidden_size * 4), nn.GELU(),
                                 nn.Linear(hidden_size * 4, hidden_size), nn.Dropout(dropout))
        self._packed = None

    def _pack(self, dev):
        ps = [self.text_proj.weight, self.image_proj.weight, self.cross_attn.in_proj_weight, self.cross_attn.out_proj.weight,
              self.mlp[0].weight, self.mlp[2].weight]
        # _version does not move when ops.FusedAdamW updates the arena in place: the optimiser generation does
        key = (tuple(p._version for p in ps), tuple(p.data_ptr() for p in ps), ops.param_generation(), str(dev))
        if self._packed is None or self._packed[0] != key:
            D = self.text_proj.out_features
            b = lambda t: t.detach().to(dev, torch.bfloat16).contiguous()
            w_in = self.cross_attn.in_proj_weight
            self._packed = (key, {"tp": b(self.text_proj.weight), "ip": b(self.image_proj.weight), "wq": b(w_in[:D]),
                                  "wkv": b(w_in[D:]), "wo": b(self.cross_attn.out_proj.weight), "w1": b(self.mlp[0].weight),
                                  "w2": b(self.mlp[2].weight)})
        return self._packed[1]

    def forward(self, hidden_states, encoder_hidden_states):
        """Inference path (eval mode / no_grad) on the tensor cores: GEMM -> LayerNorm -> single-query attention per text
        row against the projected table -> GEMM (+residual) -> LayerNorm -> MLP (+residual).

        hidden_states [B, T, text_input_size]; encoder_hidden_states [1, S, image_input_size] (what model_m.py:93-96
        passes: the vision position table, one for the whole batch).  Every text row attends to the table on its own,
        so the reference's batch-1 limitation does not apply here.  In training mode, or whenever a gradient is
        needed, the fp32 trainable path (adapter/shared_train.py) runs instead: same arithmetic, dropout, backward."""
        if encoder_hidden_states.dim() != 3 or encoder_hidden_states.shape[0] != 1:
            raise ValueError("encoder_hidden_states must be [1, S, image_input_size] (one table shared by the batch)")
        if self.training or (torch.is_grad_enabled() and (hidden_states.requires_grad or encoder_hidden_states.requires_grad
                                                          or any(p.requires_grad for p in self.parameters()))):
            # trainable path: fp32, one fused autograd node with a hand-written backward (adapter/shared_train.py);
            # dropout (adapter/clip_adapter.py:84,96) is active in training mode only
            from .shared_train import shared_adapter_train

            lead = hidden_states.shape[:-1]
            rows = hidden_states.reshape(-1, hidden_states.shape[-1]).float().contiguous()
            out = shared_adapter_train(self, rows, encoder_hidden_states[0].float().contiguous())
            return out.view(*lead, out.shape[-1])
        D, H = self.text_proj.out_features, self.cross_attn.num_heads
        if D // H != 64:
            raise ValueError("the attention kernel is specialised for head_dim = 64")
        w = self._pack(hidden_states.device)
        f = lambda t: t.detach().float()
        lead = hidden_states.shape[:-1]
        x = hidden_states.reshape(-1, hidden_states.shape[-1]).to(torch.bfloat16).contiguous()
        tab = encoder_hidden_states[0].to(torch.bfloat16).contiguous()
        S = tab.shape[0]
        h = ops.gemm(x, w["tp"], bias=f(self.text_proj.bias))
        e = ops.gemm(tab, w["ip"], bias=f(self.image_proj.bias))
        kv_in = ops.layernorm(e, f(self.norm1.weight), f(self.norm1.bias), self.norm1.eps)
        hq = ops.layernorm(h, f(self.norm2.weight), f(self.norm2.bias), self.norm2.eps)
        bias_in = f(self.cross_attn.in_proj_bias)
        q = ops.gemm(hq, w["wq"], bias=bias_in[:D].contiguous())
        kv = ops.gemm(kv_in, w["wkv"], bias=bias_in[D:].contiguous())  # [S, 2D] = [K | V]
        att = ops.attention_1q(q, kv, kv[:, D:], S, H, kv_row_stride=2 * D, kv_batch_stride=0)
        h2 = ops.gemm(att, w["wo"], bias=f(self.cross_attn.out_proj.bias), residual=hq)
        z = ops.layernorm(h2, f(self.norm3.weight), f(self.norm3.bias), self.norm3.eps)
        z = ops.gemm(z, w["w1"], bias=f(self.mlp[0].bias), act=N.ACT_GELU_ERF)
        out = ops.gemm(z, w["w2"], bias=f(self.mlp[2].bias), residual=h2)
        return out.float().view(*lead, D)

