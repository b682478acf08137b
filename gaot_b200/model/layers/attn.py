"""U-ViT processor (reference src/model/layers/attn.py:19-317) on the sm_100a kernels.

Block (attn.py:229-233, note the non-standard second residual):
    h   = x + O * SDPA(Q n1(x), K n1(x), V n1(x))
    h   = n2(h)
    out = h + W2 (silu(W1 h) * W3 h)
Q/K/V run as one GEMM on the concatenated weight, W1/W3 likewise; residual adds are GEMM epilogues."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import torch
from torch import nn

from ... import ops
from .mlp import ConditionedNorm


@dataclass
class AttentionConfig:
    num_heads: int = 8
    num_kv_heads: int = 8
    use_conditional_norm: bool = False
    cond_norm_hidden_size: int = 4
    atten_dropout: float = 0.0


@dataclass
class TransformerConfig:
    patch_size: int = 8
    hidden_size: int = 256
    use_attn_norm: bool = True
    use_ffn_norm: bool = True
    norm_eps: float = 1e-6
    num_layers: int = 3
    positional_embedding: str = "absolute"
    use_long_range_skip: bool = True
    ffn_multiplier: int = 4
    attn_config: AttentionConfig = field(default_factory=AttentionConfig)


SUPPORTED_HEAD_DIMS = (16, 32, 64)          # template instances of csrc/attn.cu and csrc/attn_tc.cu


class RMSNorm(nn.Module):
    def __init__(self, dim: int, eps: float = 1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x, mode="fp32"):
        """mode 'bf16' / 'both' additionally (or only) emit the bf16 copy a tensor-core GEMM consumes."""
        shape = x.shape
        y = ops.rmsnorm(x.reshape(-1, shape[-1]), self.weight, self.eps, mode)
        if mode == "both":
            return y[0].view(shape), y[1].view(shape)
        return y.view(shape)


class GroupQueryFlashAttention(nn.Module):
    def __init__(self, input_size: int, hidden_size: int, num_heads: int = 8, num_kv_heads: int = 8,
                 use_conditional_norm: bool = False, cond_norm_hidden_size: int = 4, atten_dropout: float = 0.0,
                 positional_embedding: str = "absolute"):
        super().__init__()
        assert hidden_size % num_heads == 0, f"hidden_size {hidden_size} must be divisible by num_heads {num_heads}"
        assert num_heads % num_kv_heads == 0, f"num_heads {num_heads} must be divisible by num_kv_heads {num_kv_heads}"
        # 'rope' (attn.py:75-76, 106-108): q and k are rotated by the token position before the attention (ops.rope_qk)
        self.use_rope = positional_embedding == "rope"
        if atten_dropout:
            raise NotImplementedError("attention dropout is not on the accelerated path")
        self.num_heads, self.num_kv_heads = num_heads, num_kv_heads
        self.num_repeat = num_heads // num_kv_heads
        self.head_dim = hidden_size // num_heads
        if self.head_dim not in SUPPORTED_HEAD_DIMS:
            raise ValueError(f"head_dim = hidden_size / num_heads = {hidden_size} / {num_heads} = {self.head_dim} is not "
                             f"instantiated in the attention kernels; supported: {SUPPORTED_HEAD_DIMS}")
        kv = self.head_dim * num_kv_heads
        self.q_proj = nn.Linear(input_size, hidden_size, bias=False)
        self.k_proj = nn.Linear(input_size, kv, bias=False)
        self.v_proj = nn.Linear(input_size, kv, bias=False)
        self.o_proj = nn.Linear(hidden_size, input_size, bias=False)
        self.correction = ConditionedNorm(1, input_size, cond_norm_hidden_size) if use_conditional_norm else None
        self._wqkv = None

    def prep_weights(self):
        """Concatenated q/k/v weight and the bf16 operands of this layer's GEMMs, ahead of time (ops.prepare_weight)."""
        self._wqkv = torch.cat([self.q_proj.weight, self.k_proj.weight, self.v_proj.weight], dim=0)
        return [self._wqkv] + ops.prepare_weight(self._wqkv) + ops.prepare_weight(self.o_proj.weight)

    @classmethod
    def from_config(cls, input_size, hidden_size, config: AttentionConfig, positional_embedding="absolute"):
        return cls(input_size, hidden_size, num_heads=config.num_heads, num_kv_heads=config.num_kv_heads,
                   use_conditional_norm=config.use_conditional_norm, cond_norm_hidden_size=config.cond_norm_hidden_size,
                   atten_dropout=config.atten_dropout, positional_embedding=positional_embedding)

    def forward(self, x, condition=None, relative_positions=None, residual=None):
        """x [B,T,D] -> o_proj(SDPA(...)) (+ residual, fused into the o_proj GEMM)."""
        B, T, D = x.shape
        if self.correction is not None:
            x = self.correction(condition, x)              # attn.py:89-90 (fp32; the GEMM below casts its operand)
        wqkv = self._wqkv if self._wqkv is not None else torch.cat([self.q_proj.weight, self.k_proj.weight, self.v_proj.weight], dim=0)
        self._wqkv = None
        tc_attn = ops.get_precision() == "bf16" and self.head_dim in (16, 32, 64) and D % 8 == 0
        qkv = ops.linear(x.reshape(B * T, D), wqkv, tc=True, out_bf16=tc_attn)
        if relative_positions is not None and self.use_rope:
            qkv = ops.rope_qk(qkv, B, T, self.num_heads, self.num_kv_heads, self.head_dim)
        o = ops.attention(qkv, B, T, self.num_heads, self.num_kv_heads, self.head_dim)
        add = residual.reshape(B * T, -1) if residual is not None else None
        return ops.linear(o, self.o_proj.weight, None, add=add, tc=True).view(B, T, -1)


class FFN(nn.Module):
    def __init__(self, input_size: int, ffn_hidden_size: int, use_conditional_norm: bool = False, cond_norm_hidden_size: int = 4):
        super().__init__()
        self.w1 = nn.Linear(input_size, ffn_hidden_size, bias=False)
        self.w2 = nn.Linear(ffn_hidden_size, input_size, bias=False)
        self.w3 = nn.Linear(input_size, ffn_hidden_size, bias=False)
        self.correction = ConditionedNorm(1, input_size, cond_norm_hidden_size) if use_conditional_norm else None
        self._w13 = None

    def _fused(self):
        """bf16 mode: the SwiGLU rides in the GEMM epilogues (ops.swiglu_ffn) when the widths fit and the output is not
        corrected before the residual (attn.py:151-154)."""
        return self.correction is None and ops.ffn_fusable(self.w1.weight.shape[1], self.w1.weight.shape[0])

    def prep_weights(self):
        if self._fused():
            self._w13 = ops.interleave_w13(self.w1.weight, self.w3.weight)
        else:
            self._w13 = torch.cat([self.w1.weight, self.w3.weight], dim=0)
        return [self._w13] + ops.prepare_weight(self._w13) + ops.prepare_weight(self.w2.weight)

    def forward(self, x, condition=None, residual=None):
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        tc16 = ops.get_precision() == "bf16" and shape[-1] % 8 == 0 and self.w1.weight.shape[0] % 8 == 0
        if self._fused():
            w13 = self._w13 if self._w13 is not None else ops.interleave_w13(self.w1.weight, self.w3.weight)
            self._w13 = None
            add = residual.reshape(-1, shape[-1]) if residual is not None else None
            return ops.swiglu_ffn(x2, w13, self.w2.weight, add=add).view(residual.shape if residual is not None else shape)
        w13 = self._w13 if self._w13 is not None else torch.cat([self.w1.weight, self.w3.weight], dim=0)
        self._w13 = None
        ab = ops.linear(x2, w13, tc=True, out_bf16=tc16)
        if self.correction is not None:
            # attn.py:151-154: the correction acts on the FFN output, so the residual cannot ride in the w2 epilogue
            y = ops.linear(ops.swiglu(ab), self.w2.weight, None, tc=True).view(shape)
            y = self.correction(condition, y)
            return y + residual if residual is not None else y
        add = residual.reshape(-1, shape[-1]) if residual is not None else None
        return ops.linear(ops.swiglu(ab), self.w2.weight, None, add=add, tc=True).view(residual.shape if residual is not None else shape)


class TransformerBlock(nn.Module):
    def __init__(self, input_size: int, config: TransformerConfig, skip_connection: bool = False):
        super().__init__()
        hidden = config.hidden_size
        self.attn = GroupQueryFlashAttention.from_config(input_size=input_size, hidden_size=hidden,
                                                         config=config.attn_config,
                                                         positional_embedding=config.positional_embedding)
        self.ffn = FFN(input_size=input_size, ffn_hidden_size=hidden * config.ffn_multiplier,
                       use_conditional_norm=config.attn_config.use_conditional_norm,
                       cond_norm_hidden_size=config.attn_config.cond_norm_hidden_size)
        self.attn_norm = RMSNorm(input_size, eps=config.norm_eps) if config.use_attn_norm else None
        self.ffn_norm = RMSNorm(input_size, eps=config.norm_eps) if config.use_ffn_norm else None
        self.skip_connection = skip_connection
        if skip_connection:
            self.skip_proj = nn.Linear(input_size * 2, input_size)

    def forward(self, x, condition=None, relative_positions=None, skip=None):
        """Positional order as the reference (attn.py:208): (x, condition, relative_positions, skip)."""
        if self.skip_connection and skip is not None:
            B, T, D = x.shape
            W = self.skip_proj.weight                                    # [D, 2D] acting on cat[x, skip] (attn.py:225-227)
            t = ops.linear(skip.reshape(B * T, D), W[:, D:], None, tc=True)
            x = ops.linear(x.reshape(B * T, D), W[:, :D], self.skip_proj.bias, add=t, tc=True).view(B, T, D)
        bf = ops.get_precision() == "bf16" and x.shape[-1] % 8 == 0 and self.attn.correction is None   # the correction wants fp32
        # bf16 mode: the norms emit the bf16 GEMM operands directly (no separate cast pass)
        h = self.attn_norm(x, "bf16" if bf else "fp32") if self.attn_norm is not None else x
        h = self.attn(h, condition=condition, relative_positions=relative_positions, residual=x)
        if self.ffn_norm is not None and bf:
            h, h16 = self.ffn_norm(h, "both")
            return self.ffn(h16, condition=condition, residual=h)
        h = self.ffn_norm(h) if self.ffn_norm is not None else h
        return self.ffn(h, condition=condition, residual=h)


class Transformer(nn.Module):
    def __init__(self, input_size: int, output_size: int, config: TransformerConfig = TransformerConfig()):
        super().__init__()
        hidden, L = config.hidden_size, config.num_layers
        self.use_long_range_skip = config.use_long_range_skip
        if input_size != hidden:
            self.input_proj = nn.Linear(input_size, hidden)
            working = hidden
        else:
            self.input_proj = nn.Identity()
            working = input_size
        self.output_proj = nn.Linear(working, output_size) if working != output_size else nn.Identity()
        self.encoder_layers = nn.ModuleList([TransformerBlock(working, config, False) for _ in range(L // 2)])
        self.middle_layer = TransformerBlock(working, config, False) if L % 2 == 1 else None
        self.decoder_layers = nn.ModuleList([TransformerBlock(working, config, True) for _ in range(L // 2)])

    def prep_weights(self):
        """bf16 GEMM operands of every block, prepared ahead of the blocks' forward (see ops.prepare_weight)."""
        out = []
        blocks = list(self.encoder_layers) + ([self.middle_layer] if self.middle_layer is not None else []) + list(self.decoder_layers)
        for blk in blocks:
            out += blk.attn.prep_weights() + blk.ffn.prep_weights()
        return out

    @staticmethod
    def _proj(layer, x):
        if isinstance(layer, nn.Identity):
            return x
        B, T, D = x.shape
        return ops.linear(x.reshape(B * T, D), layer.weight, layer.bias, tc=True).view(B, T, -1)

    def forward(self, x, condition=None, relative_positions=None):
        x = self._proj(self.input_proj, x)
        skips = []
        for blk in self.encoder_layers:
            x = blk(x, condition=condition, relative_positions=relative_positions)
            skips.append(x)
        if self.middle_layer is not None:
            x = self.middle_layer(x, condition=condition, relative_positions=relative_positions)
        for blk in self.decoder_layers:
            skip = skips.pop() if self.use_long_range_skip else None
            x = blk(x, skip=skip, condition=condition, relative_positions=relative_positions)
        return self._proj(self.output_proj, x)
