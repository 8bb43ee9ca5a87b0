"""Once-per-image cost of the two text encoders at the reference's real geometry (random weights): the native path
(visualcloze_b200.text_encoders) vs the HF bf16 modules the reference loads (models/util.py:425-431), same GPU, same ids.
Writes gpurun_out/text_encoders.json.  HF is a reference point only -- never on the product path."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformers import CLIPTextConfig, CLIPTextModel, T5Config, T5EncoderModel  # noqa: E402

from visualcloze_b200 import text_encoders as T  # noqa: E402

BF16 = torch.bfloat16


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


out = {}
cfg = T5Config(vocab_size=32128, d_model=4096, d_kv=64, d_ff=10240, num_layers=24, num_heads=64, dropout_rate=0.0,
               feed_forward_proj="gated-gelu", is_encoder_decoder=False, use_cache=False)
with torch.device("cuda"):
    hf = T5EncoderModel(cfg)
hf = hf.to(BF16).eval().requires_grad_(False)
ours = T.T5Encoder(hf.state_dict(), num_heads=64, num_layers=24)
ids = torch.randint(0, 32128, (1, 512), generator=torch.Generator().manual_seed(0)).cuda()
ids[:, 40:] = 0
with torch.no_grad():
    ref = hf(input_ids=ids, attention_mask=None).last_hidden_state
    o = ours(ids)
    out["t5_xxl_512_tokens"] = {"native_ms": timed(lambda: ours(ids)), "hf_bf16_ms": timed(lambda: hf(input_ids=ids, attention_mask=None)),
                                "rel_l2_vs_hf_bf16": rel(o, ref), "gflop": 2e-9 * 512 * 24 * (4 * 4096 * 4096 + 3 * 4096 * 10240) + 24 * 4e-9 * 512 * 512 * 4096}
from visualcloze_b200 import _lib  # noqa: E402
_lib.lib().vcb_profile_begin()
ours(ids)
cat, _ = _lib.profile_end()
out["t5_xxl_512_tokens"]["native_kernel_ms"] = {k: round(v[0], 3) for k, v in cat.items() if v[1]}
out["t5_xxl_512_tokens"]["native_launches"] = int(sum(v[1] for v in cat.values()))
print(out, flush=True)
del hf, ours
torch.cuda.empty_cache()
cfg = CLIPTextConfig(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                     max_position_embeddings=77, hidden_act="quick_gelu", eos_token_id=2, bos_token_id=49406, pad_token_id=1)
with torch.device("cuda"):
    hc = CLIPTextModel(cfg)
hc = hc.to(BF16).eval().requires_grad_(False)
oc = T.CLIPTextEncoder(hc.state_dict(), num_heads=12, num_layers=12)
ids = torch.randint(3, 49000, (1, 77), generator=torch.Generator().manual_seed(1))
ids[0, 0], ids[0, 30] = 49406, 49407
ids[0, 31:] = 1
ids = ids.cuda()
with torch.no_grad():
    r = hc(input_ids=ids, attention_mask=None)
    last, pooled = oc(ids)
    out["clip_l_77_tokens"] = {"native_ms": timed(lambda: oc(ids)), "hf_bf16_ms": timed(lambda: hc(input_ids=ids, attention_mask=None)),
                               "rel_l2_pooled_vs_hf_bf16": rel(pooled, r.pooler_output)}
print(out)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/text_encoders.json", "w"), indent=1)
