"""TEST INFRASTRUCTURE ONLY — stand-in for diffusers 0.21.2 `Attention` (requirements.txt:2 of the reference).

Mirrors the semantics hook.py:92-120 relies on (written from the published diffusers 0.21.2
behaviour; diffusers itself is not available offline):
  head_to_batch_dim : [B,S,H*d] -> [B*H,S,d], batch-major (row = b*H + head)
  batch_to_head_dim : inverse
  get_attention_scores : baddbmm(alpha=scale) -> softmax(dim=-1) -> cast back to the input dtype
  to_q/to_k/to_v bias-free Linear; to_out = [Linear(bias), Dropout]; scale = d**-0.5
"""
import torch
from torch import nn


class Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0, bias=False,
                 upcast_attention=False, upcast_softmax=False):
        super().__init__()
        inner_dim = dim_head * heads
        cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.upcast_attention = upcast_attention
        self.upcast_softmax = upcast_softmax
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.norm_cross = None
        self.to_q = nn.Linear(query_dim, inner_dim, bias=bias)
        self.to_k = nn.Linear(cross_attention_dim, inner_dim, bias=bias)
        self.to_v = nn.Linear(cross_attention_dim, inner_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner_dim, query_dim), nn.Dropout(dropout)])
        self.processor = None

    def set_processor(self, processor):
        self.processor = processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask)

    def prepare_attention_mask(self, attention_mask, target_length, batch_size):
        if attention_mask is None:
            return None
        raise NotImplementedError("the SD UNet passes no attention mask on this path")

    def head_to_batch_dim(self, tensor):
        b, s, c = tensor.shape
        h = self.heads
        return tensor.reshape(b, s, h, c // h).permute(0, 2, 1, 3).reshape(b * h, s, c // h)

    def batch_to_head_dim(self, tensor):
        bh, s, d = tensor.shape
        h = self.heads
        return tensor.reshape(bh // h, h, s, d).permute(0, 2, 1, 3).reshape(bh // h, s, d * h)

    def get_attention_scores(self, query, key, attention_mask=None):
        dtype = query.dtype
        if self.upcast_attention:
            query, key = query.float(), key.float()
        if attention_mask is None:
            base = torch.empty(query.shape[0], query.shape[1], key.shape[1], dtype=query.dtype,
                               device=query.device)
            beta = 0
        else:
            base, beta = attention_mask, 1
        scores = torch.baddbmm(base, query, key.transpose(-1, -2), beta=beta, alpha=self.scale)
        del base
        if self.upcast_softmax:
            scores = scores.float()
        probs = scores.softmax(dim=-1)
        del scores
        return probs.to(dtype)
