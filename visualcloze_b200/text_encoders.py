"""The two text encoders on the far side of the hot path, native: the reference wraps HF ``T5EncoderModel``
("google/t5-v1_1-xxl", 512 tokens -> ``txt`` [B, 512, 4096]) and ``CLIPTextModel`` ("openai/clip-vit-large-patch14", 77 tokens ->
pooled ``y`` [B, 768]) in ``HFEmbedder`` (models/modules/conditioner.py:5-37, bf16 weights, ``attention_mask=None``) and calls
them once per image from ``prepare_modified`` (models/sampling.py:91-105).

Here every Linear is one libvcb200 GEMM launch (fused q|k|v and wi_0|wi_1 weights, residual adds in the GEMM epilogue) and the
rest are the small kernels of ``csrc/text_kernels.cuh``; state-dict keys are HF's, so the same checkpoints load.  Tokenisation
stays with the HF tokenizers (CPU string processing, out of scope).  ``HFEmbedder`` below mirrors the reference class: it takes
text and returns what ``prepare_modified`` consumes.  Rounding points follow the bf16 HF modules; parity is tolerance-based
(tests/test_text_encoders_gpu.py) because HF's own bf16 elementwise chains (NewGELU, SDPA) are not bit-reproducible across
backends either."""
from __future__ import annotations

import math

import torch
from torch import Tensor

from . import ops

BF16 = torch.bfloat16


def t5_relative_position_bucket(relative_position: Tensor, num_buckets: int = 32, max_distance: int = 128) -> Tensor:
    """Bidirectional bucket of ``memory_position - query_position`` (HF modeling_t5.T5Attention._relative_position_bucket,
    the Mesh-TensorFlow rule): half of the buckets per sign; of those, half exact, half logarithmic up to max_distance."""
    nb = num_buckets // 2
    buckets = (relative_position > 0).to(torch.long) * nb
    rp = relative_position.abs()
    max_exact = nb // 2
    is_small = rp < max_exact
    large = max_exact + (torch.log(rp.float() / max_exact) / math.log(max_distance / max_exact) * (nb - max_exact)).to(torch.long)
    large = torch.min(large, torch.full_like(large, nb - 1))
    return buckets + torch.where(is_small, rp, large)


class _Packed:
    """parameter packing shared by both encoders: bf16 device tensors, fp32 biases (what vcb_gemm_bf16 takes)"""

    def __init__(self, state_dict: dict, device):
        self.sd, self.device = state_dict, torch.device(device)
        if self.device.type != "cuda":
            from ._lib import VcbError
            raise VcbError("the text encoders need a CUDA device: there is no CPU fallback")

    def w(self, *names: str) -> Tensor:
        """one weight, or several concatenated along the output dimension (fused projections)"""
        ts = [self.sd[n] for n in names]
        return torch.cat([t.to(self.device, BF16) for t in ts], dim=0).contiguous()

    def b(self, *names: str) -> Tensor:
        return torch.cat([self.sd[n].to(self.device, torch.float32) for n in names], dim=0).contiguous()

    def v(self, name: str) -> Tensor:
        return self.sd[name].to(self.device, BF16).contiguous()


class T5Encoder:
    """HF ``T5EncoderModel`` forward (encoder stack of T5 v1.1: pre-RMSNorm blocks, relative-position-bias self-attention without
    1/sqrt(d) scaling, gated-GELU feed-forward, final RMSNorm), ``attention_mask=None`` as the reference calls it."""

    def __init__(self, state_dict: dict, *, num_heads: int, num_layers: int, d_kv: int = 64, num_buckets: int = 32,
                 max_distance: int = 128, eps: float = 1e-6, device="cuda"):
        if d_kv != 64:
            raise ValueError("the attention kernel serves head_dim 64 (T5 v1.1 and CLIP-L)")
        P = _Packed(state_dict, device)
        self.device, self.heads, self.layers, self.eps = P.device, num_heads, num_layers, eps
        self.num_buckets, self.max_distance = num_buckets, max_distance
        self.embed = P.v("shared.weight")
        self.d_model = self.embed.shape[1]
        self.inner = num_heads * 64
        self.blocks = []
        for i in range(num_layers):
            a, f = f"encoder.block.{i}.layer.0.", f"encoder.block.{i}.layer.1."
            self.blocks.append(dict(
                ln1=P.v(a + "layer_norm.weight"),
                qkv=P.w(a + "SelfAttention.q.weight", a + "SelfAttention.k.weight", a + "SelfAttention.v.weight"),
                o=P.w(a + "SelfAttention.o.weight"),
                ln2=P.v(f + "layer_norm.weight"),
                wi=P.w(f + "DenseReluDense.wi_0.weight", f + "DenseReluDense.wi_1.weight"),
                wo=P.w(f + "DenseReluDense.wo.weight")))
        self.d_ff = self.blocks[0]["wo"].shape[1] if self.blocks else 0
        self.final_ln = P.v("encoder.final_layer_norm.weight")
        self.rel_table = P.sd["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"].to(P.device, BF16)   # [buckets, heads]
        self._bias_cache: dict[int, Tensor] = {}

    def position_bias(self, L: int) -> Tensor:
        """[heads, L, L] bf16: T5Attention.compute_bias, shared by every layer"""
        if L not in self._bias_cache:
            ctx = torch.arange(L, device=self.device)[:, None]
            mem = torch.arange(L, device=self.device)[None, :]
            bucket = t5_relative_position_bucket(mem - ctx, self.num_buckets, self.max_distance)
            self._bias_cache[L] = self.rel_table[bucket].permute(2, 0, 1).contiguous()
        return self._bias_cache[L]

    @torch.no_grad()
    def __call__(self, input_ids: Tensor) -> Tensor:
        B, L = input_ids.shape
        n, d, dev = B * L, self.d_model, self.device
        ids = input_ids.to(dev, torch.int64).reshape(-1).contiguous()
        h = torch.empty(n, d, dtype=BF16, device=dev)
        ops.embedding(self.embed, ids, h)
        xn = torch.empty(n, d, dtype=BF16, device=dev)
        qkv = torch.empty(n, 3 * self.inner, dtype=BF16, device=dev)
        att = torch.empty(n, self.inner, dtype=BF16, device=dev)
        ab = torch.empty(n, 2 * self.d_ff, dtype=BF16, device=dev)
        ff = torch.empty(n, self.d_ff, dtype=BF16, device=dev)
        bias = self.position_bias(L)
        I = self.inner
        for blk in self.blocks:
            ops.rmsnorm_weight(h, blk["ln1"], xn, self.eps)
            ops.gemm(xn, blk["qkv"], None, qkv)
            ops.attention_small(qkv[:, :I], qkv[:, I:2 * I], qkv[:, 2 * I:], att, B, L, self.heads, bias=bias)
            ops.gemm(att, blk["o"], None, h, epilogue=ops.EPI_GATE_RES, res=h)            # h += o(att)
            ops.rmsnorm_weight(h, blk["ln2"], xn, self.eps)
            ops.gemm(xn, blk["wi"], None, ab)
            ops.gated_gelu(ab, ff)
            ops.gemm(ff, blk["wo"], None, h, epilogue=ops.EPI_GATE_RES, res=h)            # h += wo(gelu(wi_0 x) * wi_1 x)
        out = torch.empty(n, d, dtype=BF16, device=dev)
        ops.rmsnorm_weight(h, self.final_ln, out, self.eps)
        return out.reshape(B, L, d)


class CLIPTextEncoder:
    """HF ``CLIPTextModel`` forward: token + position embeddings, pre-LayerNorm blocks with causal attention and quick-GELU MLP,
    final LayerNorm; returns (last_hidden_state, pooler_output) -- the reference reads ``pooler_output`` (conditioner.py:10)."""

    def __init__(self, state_dict: dict, *, num_heads: int, num_layers: int, eps: float = 1e-5, eos_token_id: int = 2, device="cuda"):
        P = _Packed(state_dict, device)
        self.device, self.heads, self.layers, self.eps, self.eos_token_id = P.device, num_heads, num_layers, eps, eos_token_id
        t = "text_model."
        self.tok = P.v(t + "embeddings.token_embedding.weight")
        self.pos = P.v(t + "embeddings.position_embedding.weight")
        self.d = self.tok.shape[1]
        if self.d != num_heads * 64:
            raise ValueError("the attention kernel serves head_dim 64 (hidden_size == 64 * num_heads)")
        self.blocks = []
        for i in range(num_layers):
            p = f"{t}encoder.layers.{i}."
            self.blocks.append(dict(
                ln1w=P.v(p + "layer_norm1.weight"), ln1b=P.v(p + "layer_norm1.bias"),
                qkv=P.w(p + "self_attn.q_proj.weight", p + "self_attn.k_proj.weight", p + "self_attn.v_proj.weight"),
                qkv_b=P.b(p + "self_attn.q_proj.bias", p + "self_attn.k_proj.bias", p + "self_attn.v_proj.bias"),
                o=P.w(p + "self_attn.out_proj.weight"), o_b=P.b(p + "self_attn.out_proj.bias"),
                ln2w=P.v(p + "layer_norm2.weight"), ln2b=P.v(p + "layer_norm2.bias"),
                fc1=P.w(p + "mlp.fc1.weight"), fc1_b=P.b(p + "mlp.fc1.bias"),
                fc2=P.w(p + "mlp.fc2.weight"), fc2_b=P.b(p + "mlp.fc2.bias")))
        self.d_ff = self.blocks[0]["fc1"].shape[0] if self.blocks else 0
        self.flnw, self.flnb = P.v(t + "final_layer_norm.weight"), P.v(t + "final_layer_norm.bias")

    @torch.no_grad()
    def __call__(self, input_ids: Tensor) -> tuple[Tensor, Tensor]:
        B, L = input_ids.shape
        if L > self.pos.shape[0]:
            raise ValueError(f"sequence length {L} exceeds the {self.pos.shape[0]} position embeddings")
        n, d, dev = B * L, self.d, self.device
        ids2 = input_ids.to(dev, torch.int64)
        ids = ids2.reshape(-1).contiguous()
        h = torch.empty(n, d, dtype=BF16, device=dev)
        ops.embedding(self.tok, ids, h, pos_table=self.pos, L=L)
        xn = torch.empty(n, d, dtype=BF16, device=dev)
        qkv = torch.empty(n, 3 * d, dtype=BF16, device=dev)
        att = torch.empty(n, d, dtype=BF16, device=dev)
        f1 = torch.empty(n, self.d_ff, dtype=BF16, device=dev)
        for blk in self.blocks:
            ops.layernorm_affine(h, blk["ln1w"], blk["ln1b"], xn, self.eps)
            ops.gemm(xn, blk["qkv"], blk["qkv_b"], qkv)
            ops.attention_small(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], att, B, L, self.heads, scale=0.125, causal=True)
            ops.gemm(att, blk["o"], blk["o_b"], h, epilogue=ops.EPI_GATE_RES, res=h)
            ops.layernorm_affine(h, blk["ln2w"], blk["ln2b"], xn, self.eps)
            ops.gemm(xn, blk["fc1"], blk["fc1_b"], f1)
            ops.quick_gelu(f1, f1)
            ops.gemm(f1, blk["fc2"], blk["fc2_b"], h, epilogue=ops.EPI_GATE_RES, res=h)
        last = torch.empty(n, d, dtype=BF16, device=dev)
        ops.layernorm_affine(h, self.flnw, self.flnb, last, self.eps)
        last = last.reshape(B, L, d)
        if self.eos_token_id == 2:            # the legacy rule of the openai/clip-vit-large-patch14 config: EOS has the largest id
            eos = ids2.argmax(dim=-1)
        else:
            eos = (ids2 == self.eos_token_id).int().argmax(dim=-1)
        pooled = last[torch.arange(B, device=dev), eos]
        return last, pooled


class HFEmbedder:
    """The reference's ``HFEmbedder`` (models/modules/conditioner.py:5-37) over the native encoders: ``forward(text: list[str])``
    tokenises with the HF tokenizer (``padding="max_length"``, truncation) and returns ``pooler_output`` for CLIP and
    ``last_hidden_state`` for T5 -- what ``prepare_modified`` consumes."""

    def __init__(self, tokenizer, encoder, max_length: int):
        self.tokenizer, self.encoder, self.max_length = tokenizer, encoder, max_length
        self.is_clip = isinstance(encoder, CLIPTextEncoder)

    def __call__(self, text: list[str]) -> Tensor:
        enc = self.tokenizer(text, truncation=True, max_length=self.max_length, return_length=False, return_overflowing_tokens=False,
                             padding="max_length", return_tensors="pt")
        out = self.encoder(enc["input_ids"])
        return out[1] if self.is_clip else out

    forward = __call__
