"""Import stand-in for the `rotary-embedding-torch` wheel (lucidrains; listed UNPINNED in the reference's requirements.txt:11,
absent from this image).  Restates the part of its published algorithm the reference calls - `RotaryEmbedding(dim=head_dim)` and
`rotate_queries_or_keys(t)` on [batch, heads, seq, head_dim] (attn.py:76, 106-108) - with the library defaults: theta = 10000,
language-model frequencies, no learned frequencies, no xpos scaling, interpolation factor 1, positions 0..seq-1 along dim -2.
Parity of this option is therefore pinned to this restatement, not to the wheel ("parity unpinned")."""
import torch


class RotaryEmbedding(torch.nn.Module):
    def __init__(self, dim, theta=10000):
        super().__init__()
        self.dim = dim
        self.register_buffer("freqs", 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim)), persistent=False)

    def rotate_queries_or_keys(self, t, seq_dim=-2):
        seq = torch.arange(t.shape[seq_dim], device=t.device, dtype=self.freqs.dtype)
        ang = torch.einsum("i,j->ij", seq, self.freqs.to(t.device))           # [seq, dim / 2]
        ang = ang.repeat_interleave(2, dim=-1).to(t.dtype)                        # (f0, f0, f1, f1, ...)
        x1, x2 = t[..., 0::2], t[..., 1::2]
        rot = torch.stack((-x2, x1), dim=-1).reshape(t.shape)                     # rotate_half on interleaved pairs
        return t * ang.cos() + rot * ang.sin()
