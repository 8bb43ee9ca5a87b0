"""The native text encoders against the modules the reference loads: HF ``T5EncoderModel`` and ``CLIPTextModel`` in
bf16, called with ``attention_mask=None`` (models/modules/conditioner.py:5-37).  Weights are random (no network for checkpoints);
the HF modules are built from configs of the real geometry and share their state dict with ours.

Kernel level: each helper against the torch expression HF evaluates (same rounding points).  Model level: hidden states against HF
in bf16 on the same GPU, read against HF's own bf16-vs-fp32 distance -- two bf16 evaluation orders of a deep residual stack differ
by about that much, and HF's bf16 elementwise chains (NewGELU as six rounded ops, SDPA) are not what we reproduce bit for bit.
"""
import math

import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from visualcloze_b200 import ops as o
    return o


def test_rmsnorm_weight_matches_t5_layernorm(ops):
    from transformers.models.t5.modeling_t5 import T5LayerNorm
    g = torch.Generator(device="cuda").manual_seed(1)
    x = (torch.randn(777, 4096, generator=g, device="cuda") * 3).to(BF16)
    ln = T5LayerNorm(4096, eps=1e-6).to("cuda", BF16)
    ln.weight.data = (1 + 0.2 * torch.randn(4096, generator=g, device="cuda")).to(BF16)
    ref = ln(x)
    out = torch.empty_like(x)
    ops.rmsnorm_weight(x, ln.weight.data, out, 1e-6)
    torch.cuda.synchronize()
    assert out.dtype == BF16 and rel_l2(out.float(), ref.float()) < 1e-3
    assert float((out.float() - ref.float()).abs().max()) <= 0.0625            # a bf16 ulp at |y| <= 8: rsqrtf vs torch.rsqrt


def test_layernorm_affine_matches_torch(ops):
    g = torch.Generator(device="cuda").manual_seed(2)
    x = (torch.randn(154, 768, generator=g, device="cuda") * 2 + 0.5).to(BF16)
    w = (1 + 0.2 * torch.randn(768, generator=g, device="cuda")).to(BF16)
    b = (0.3 * torch.randn(768, generator=g, device="cuda")).to(BF16)
    ref = torch.nn.functional.layer_norm(x, (768,), w, b, 1e-5)
    out = torch.empty_like(x)
    ops.layernorm_affine(x, w, b, out, 1e-5)
    torch.cuda.synchronize()
    assert rel_l2(out.float(), ref.float()) < 2e-3


def test_gated_gelu_and_quick_gelu(ops):
    from transformers.activations import NewGELUActivation
    g = torch.Generator(device="cuda").manual_seed(3)
    ab = (torch.randn(300, 2 * 1024, generator=g, device="cuda") * 2).to(BF16)
    ref = NewGELUActivation()(ab[:, :1024]) * ab[:, 1024:]
    out = torch.empty(300, 1024, dtype=BF16, device="cuda")
    ops.gated_gelu(ab, out)
    x = (torch.randn(77, 3072, generator=g, device="cuda") * 2).to(BF16)
    q = torch.empty_like(x)
    ops.quick_gelu(x, q)
    torch.cuda.synchronize()
    assert rel_l2(out.float(), ref.float()) < 1e-2           # HF rounds each of the six elementwise ops of NewGELU to bf16
    truth = (torch.nn.functional.gelu(ab[:, :1024].float(), approximate="tanh").to(BF16).float() * ab[:, 1024:].float())
    assert rel_l2(out.float(), truth) < 4e-3                 # against the function itself: one rounding of the gelu, one of the product
    assert rel_l2(q.float(), x.float() * torch.sigmoid(1.702 * x.float())) < 4e-3


def _attn_ref(q, k, v, B, L, heads, bias, scale, causal):
    """the HF rounding points in torch: bf16 scores (x scale), bf16 bias add, fp32 softmax, bf16 probabilities, bf16 output"""
    qh, kh, vh = (t.reshape(B, L, heads, 64).permute(0, 2, 1, 3).float() for t in (q, k, v))
    s = (qh @ kh.transpose(-1, -2)).to(BF16)
    if scale != 1.0:
        s = (s.float() * scale).to(BF16)
    if bias is not None:
        s = (s.float() + bias.float()[None]).to(BF16)
    s = s.float()
    if causal:
        s = s.masked_fill(torch.triu(torch.ones(L, L, dtype=torch.bool, device=s.device), 1), float("-inf"))
    p = torch.softmax(s, dim=-1).to(BF16).float()
    return (p @ vh).to(BF16).permute(0, 2, 1, 3).reshape(B * L, heads * 64)


@pytest.mark.parametrize("case", [(2, 512, 4, True, 1.0, False), (1, 100, 3, True, 1.0, False), (3, 77, 12, False, 0.125, True), (1, 16, 2, False, 0.125, True)])
def test_attention_small(ops, case):
    B, L, heads, with_bias, scale, causal = case
    g = torch.Generator(device="cuda").manual_seed(L)
    qkv = torch.randn(B * L, 3 * heads * 64, generator=g, device="cuda").to(BF16)
    if with_bias:
        qkv[:, :heads * 64] *= 0.3                 # T5 has no 1/sqrt(d): keep the logits in a trained model's range
    bias = (2 * torch.randn(heads, L, L, generator=g, device="cuda")).to(BF16) if with_bias else None
    I = heads * 64
    out = torch.full((B * L, I), 7.0, dtype=BF16, device="cuda")
    ops.attention_small(qkv[:, :I], qkv[:, I:2 * I], qkv[:, 2 * I:], out, B, L, heads, bias=bias, scale=scale, causal=causal)
    torch.cuda.synchronize()
    ref = _attn_ref(qkv[:, :I], qkv[:, I:2 * I], qkv[:, 2 * I:], B, L, heads, bias, scale, causal)
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out.float(), ref.float()) < 6e-3
    with pytest.raises(Exception, match="attention_small"):
        ops.attention_small(qkv[:, :I], qkv[:, I:2 * I], qkv[:, 2 * I:], out, B, 513, heads)


def _t5_pair(d_model, heads, d_ff, layers, vocab, seed):
    from transformers import T5Config, T5EncoderModel
    from visualcloze_b200 import text_encoders as T
    cfg = T5Config(vocab_size=vocab, d_model=d_model, d_kv=64, d_ff=d_ff, num_layers=layers, num_heads=heads,
                   relative_attention_num_buckets=32, relative_attention_max_distance=128, dropout_rate=0.0,
                   layer_norm_epsilon=1e-6, feed_forward_proj="gated-gelu", is_encoder_decoder=False, use_cache=False)
    torch.manual_seed(seed)
    with torch.device("cuda"):
        hf = T5EncoderModel(cfg)
    hf = hf.eval().requires_grad_(False)
    g = torch.Generator(device="cuda").manual_seed(seed)
    for n, p in hf.named_parameters():          # HF's init leaves the norms at 1 and the bias table small: make every tensor matter
        if n.endswith("layer_norm.weight"):
            p.data = 1 + 0.1 * torch.randn(p.shape, generator=g, device="cuda")
        elif "relative_attention_bias" in n:
            p.data = torch.randn(p.shape, generator=g, device="cuda")
    hf16 = T5EncoderModel(cfg).to("cuda", BF16).eval().requires_grad_(False)
    hf16.load_state_dict({k: v.to(BF16) for k, v in hf.state_dict().items()})
    ours = T.T5Encoder(hf16.state_dict(), num_heads=heads, num_layers=layers)
    return hf, hf16, ours


@pytest.mark.parametrize("geom", [(256, 4, 512, 3, 1000, 2, 128), (4096, 64, 10240, 2, 32128, 1, 512)])
def test_t5_encoder_vs_hf(geom):
    """small geometry (3 layers) and the XXL geometry of google/t5-v1_1-xxl (d_model 4096, 64 heads, d_ff 10240; 2 of its 24
    layers, 512 tokens as the pipeline passes them)"""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    d_model, heads, d_ff, layers, vocab, B, L = geom
    hf32, hf16, ours = _t5_pair(d_model, heads, d_ff, layers, vocab, seed=d_model)
    ids = torch.randint(0, vocab, (B, L), generator=torch.Generator().manual_seed(5)).cuda()
    ids[:, L // 2:] = 0                                       # the pad tail the tokenizer produces (attended: attention_mask=None)
    with torch.no_grad():
        ref16 = hf16(input_ids=ids, attention_mask=None).last_hidden_state.float()
        ref32 = hf32(input_ids=ids, attention_mask=None).last_hidden_state.float()
    out = ours(ids)
    torch.cuda.synchronize()
    assert out.shape == (B, L, d_model) and out.dtype == BF16 and torch.isfinite(out.float()).all()
    e16, e32, floor = rel_l2(out.float(), ref16), rel_l2(out.float(), ref32), rel_l2(ref16, ref32)
    print(f"[t5 {d_model}x{layers}] ours vs HF bf16 {e16:.3e}, ours vs HF fp32 {e32:.3e}, HF bf16 vs HF fp32 {floor:.3e}")
    assert e16 < 2e-2, (e16, e32, floor)
    assert e32 < 1.5 * floor + 2e-3, (e16, e32, floor)


@pytest.mark.parametrize("eos", [2, 49407])
def test_clip_text_encoder_vs_hf(eos):
    """the openai/clip-vit-large-patch14 text geometry (768 wide, 12 heads, 3072 MLP, 77 positions; 3 of its 12 layers), both
    pooling rules: the checkpoint's legacy eos_token_id = 2 (argmax of the ids) and an explicit EOS id"""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from transformers import CLIPTextConfig, CLIPTextModel
    from visualcloze_b200 import text_encoders as T
    cfg = CLIPTextConfig(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=3, num_attention_heads=12,
                         max_position_embeddings=77, hidden_act="quick_gelu", layer_norm_eps=1e-5, eos_token_id=eos, bos_token_id=49406,
                         pad_token_id=1)
    torch.manual_seed(7)
    with torch.device("cuda"):
        hf32 = CLIPTextModel(cfg)
    hf32 = hf32.eval().requires_grad_(False)
    g = torch.Generator(device="cuda").manual_seed(7)
    for n, p in hf32.named_parameters():
        if "layer_norm" in n or n.endswith(".bias"):
            p.data = (1.0 if n.endswith("norm1.weight") or n.endswith("norm2.weight") or n.endswith("final_layer_norm.weight") else 0.0) \
                + 0.1 * torch.randn(p.shape, generator=g, device="cuda")
        elif p.dim() == 2 and "embedding" not in n:
            p.data = torch.randn(p.shape, generator=g, device="cuda") / math.sqrt(p.shape[1])
    hf16 = CLIPTextModel(cfg).to("cuda", BF16).eval().requires_grad_(False)
    hf16.load_state_dict({k: v.to(BF16) for k, v in hf32.state_dict().items()})
    ours = T.CLIPTextEncoder(hf16.state_dict(), num_heads=12, num_layers=3, eos_token_id=eos)
    B, L = 2, 77
    ids = torch.randint(3, 49000, (B, L), generator=torch.Generator().manual_seed(9))
    ids[:, 0] = 49406
    ids[0, 20], ids[1, 41] = 49407, 49407                      # EOS: the largest id, so both pooling rules pick it
    ids[0, 21:], ids[1, 42:] = 1, 1
    ids = ids.cuda()
    with torch.no_grad():
        r16 = hf16(input_ids=ids, attention_mask=None)
        r32 = hf32(input_ids=ids, attention_mask=None)
    last, pooled = ours(ids)
    torch.cuda.synchronize()
    assert last.shape == (B, L, 768) and pooled.shape == (B, 768)
    e_last, floor = rel_l2(last.float(), r16.last_hidden_state.float()), rel_l2(r16.last_hidden_state.float(), r32.last_hidden_state.float())
    e_pool = rel_l2(pooled.float(), r16.pooler_output.float())
    print(f"[clip eos={eos}] last vs HF bf16 {e_last:.3e} (HF bf16 vs fp32 {floor:.3e}), pooled vs HF bf16 {e_pool:.3e}")
    assert e_last < 2e-2 and e_pool < 2e-2, (e_last, e_pool, floor)
    assert rel_l2(last.float(), r32.last_hidden_state.float()) < 1.5 * floor + 2e-3
    assert torch.equal(pooled, last[torch.arange(B, device="cuda"), torch.tensor([20, 41], device="cuda")])
