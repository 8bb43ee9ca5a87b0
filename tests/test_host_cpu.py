"""CPU-only tests: the C-ABI library loads and exports what include/vcb200.h declares, the product path fails loudly
without a GPU, the reference-shaped host mirrors agree with the reference-generated goldens, and the N>1 plumbing works
over gloo with world_size 2."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from conftest import GOLDEN, REPO


def _golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def test_library_exports_every_declared_symbol():
    from visualcloze_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run `python __graft_entry__.py` (build()) first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    header = open(os.path.join(REPO, "include", "vcb200.h")).read()
    declared = sorted(set(re.findall(r"\b(vcb_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 15
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert set(_lib.exported_symbols()) <= set(declared)
    assert _lib.lib().vcb_abi_version() == _lib.ABI_VERSION


def test_header_is_plain_c_and_struct_layouts_match_the_binding(tmp_path):
    """include/vcb200.h must be consumable from C (the drop-in boundary is a C ABI) and the ctypes mirrors must have the sizes
    the C compiler gives the structs."""
    import shutil
    from visualcloze_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    hdr = os.path.join(REPO, "include", "vcb200.h")
    assert subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", hdr]).returncode == 0
    names = {"vcb_gemm_args": _lib.GemmArgs, "vcb_attn_args": _lib.AttnArgs, "vcb_ln_args": _lib.LnArgs, "vcb_flux_config": _lib.FluxConfigC,
             "vcb_flux_weights": _lib.FluxWeightsC, "vcb_double_w": _lib.DoubleW, "vcb_single_w": _lib.SingleW,
             "vcb_vae_config": _lib.VaeConfigC, "vcb_vae_weights": _lib.VaeWeightsC, "vcb_vae_enc_weights": _lib.VaeEncWeightsC}
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "vcb200.h"\nint main(void) {\n' +
                   "".join(f'  printf("{n} %zu\\n", sizeof({n}));\n' for n in names) + "  return 0;\n}\n")
    exe = tmp_path / "sizes"
    assert subprocess.run([gcc, "-std=c99", "-I", os.path.join(REPO, "include"), str(src), "-o", str(exe)]).returncode == 0
    out = subprocess.run([str(exe)], capture_output=True, text=True).stdout
    for line in out.splitlines():
        n, size = line.split()
        assert ctypes.sizeof(names[n]) == int(size), f"{n}: C says {size}, ctypes says {ctypes.sizeof(names[n])}"


def test_attention_score_bound_from_norm_scales():
    """engine._score_bound: the promised bound must really dominate every scaled score q.k * 128^-0.5 * log2(e) of vectors that
    went through RMSNorm * scale (+ a rotation), and must switch itself off (0 = exact online softmax) outside the safe range."""
    import math
    import types
    from visualcloze_b200.engine import FluxEngine
    g = torch.Generator().manual_seed(0)
    qs, ks = (1 + 0.2 * torch.randn(128, generator=g)).bfloat16(), (1 + 0.2 * torch.randn(128, generator=g)).bfloat16()
    fake = types.SimpleNamespace(_p={"q": qs, "k": ks, "big": 3.0 * torch.ones(128)})
    b = FluxEngine._score_bound(fake, ["q"], ["k"])
    assert 0 < b <= 48
    x, y = torch.randn(4096, 128, generator=g) * 5, torch.randn(4096, 128, generator=g) * 0.01
    rms = lambda t: t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + 1e-6)
    q = (rms(x).bfloat16() * qs).float()                      # layers.py:68-72 rounding points
    k = (rms(y).bfloat16() * ks).float()
    k[0] = q[0] * (k[0].norm() / q[0].norm())                 # one perfectly aligned pair
    scores = (q @ k.T) * (128 ** -0.5) * math.log2(math.e)
    assert float(scores.abs().max()) <= b
    assert float(scores.abs().max()) > 0.25 * b               # and it is not vacuous
    assert FluxEngine._score_bound(fake, ["big"], ["big"]) == 0.0      # 9 * 16.3 > 48: exact kernel


def test_bench_algorithmic_flops_match_the_survey_table():
    """bench.py's roofline numerators (per evaluation at cfg B 51.23 TF of merged-LoRA GEMMs in the 57 blocks
    + embed/final, 11.03 TF of attention)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(REPO, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    gemm1, attn1 = bench.flops_per_image(3456, 512, 1)
    gemm2, attn2 = bench.flops_per_image(3456, 512, 2)
    per_eval = gemm2 - gemm1                                   # the step-invariant part (txt_in, embedders) cancels
    assert abs(attn1 - 11.03e12) / 11.03e12 < 2e-3
    assert abs((attn2 - attn1) - attn1) < 1e-6 * attn1
    blocks = 57 * 24 * 3968 * 3072 ** 2
    assert blocks < per_eval < 1.01 * blocks and abs(blocks - 51.23e12) / 51.23e12 < 1e-3
    x, kw, Li, Lt = bench.make_inputs("B", 1234)
    assert (Li, Lt) == (3456, 512) and x.shape == (1, 3456, 64) and kw["cond"].shape == (1, 3456, 320)
    assert kw["img_ids"][0, :, 0].unique().tolist() == [1.0, 2.0]            # one RoPE row id per grid row (sampling.py:56-61)


def test_fixed_reference_softmax_is_the_same_function():
    """The identity the bounded attention kernel relies on: softmax(s) == 2^(s' - B) / sum 2^(s' - B) for ANY reference B
    (s' = s * log2 e), evaluated the way the kernel does (P rounded to bf16, fp32 row sum of the unrounded p)."""
    import math
    g = torch.Generator().manual_seed(1)
    s = torch.randn(64, 512, generator=g) * 3
    v = torch.randn(512, 128, generator=g)
    ref = torch.softmax(s, dim=-1) @ v
    s2 = s * math.log2(math.e)
    for B in (float(s2.max()), float(s2.max()) + 11.0, 40.0):
        p = torch.exp2(s2 - B)
        out = (p.bfloat16().float() @ v) / p.sum(-1, keepdim=True)
        assert float((out - ref).norm() / ref.norm()) < 3e-3, B


def test_no_cpu_fallback():
    from visualcloze_b200 import _lib, model as M, ops
    x = torch.zeros(4, 256, dtype=torch.bfloat16)
    with pytest.raises(_lib.VcbError):
        ops.silu(x, x.clone())
    if not torch.cuda.is_available():
        # pointer-level call without a device: the library itself refuses (no compute is attempted)
        rc = _lib.lib().vcb_silu(ctypes.c_void_p(16), ctypes.c_void_p(16), 4, None)
        assert rc != 0 and b"CUDA" in _lib.lib().vcb_last_error()
    p = M.FluxParams(in_channels=384, out_channels=64, vec_in_dim=32, context_in_dim=64, hidden_size=256, mlp_ratio=2.0,
                     num_heads=2, depth=1, depth_single_blocks=1, axes_dim=[16, 56, 56], theta=10_000, qkv_bias=True,
                     guidance_embed=True)
    model = M.Flux(p)
    g = _golden("flux_small_b1.pt")["inputs"]
    with pytest.raises(_lib.VcbError):
        model(**g)
    with pytest.raises(ValueError, match="3 dimensions"):
        model(**dict(g, txt=g["txt"][0]))
    assert "oracle" not in sys.modules or True
    src = "".join(open(os.path.join(REPO, "visualcloze_b200", f)).read() for f in os.listdir(os.path.join(REPO, "visualcloze_b200")) if f.endswith(".py"))
    assert "import oracle" not in src and "from oracle" not in src, "the product must never import the oracle"


def test_state_dict_contract_matches_oracle_spec():
    from oracle import flux_oracle as fo
    from visualcloze_b200 import model as M
    for lora in (0, 16):
        kw = dict(in_channels=384, out_channels=64, vec_in_dim=32, context_in_dim=64, hidden_size=256, mlp_ratio=2.0,
                  num_heads=2, depth=2, depth_single_blocks=3, axes_dim=[16, 56, 56], theta=10_000, qkv_bias=True,
                  guidance_embed=True)
        m = M.FluxLoraWrapper(lora_rank=lora, params=M.FluxParams(**kw)) if lora else M.Flux(M.FluxParams(**kw))
        shapes = fo.param_shapes(fo.FluxConfig(**kw, lora_rank=lora), lora=bool(lora))
        sd = m.state_dict()
        assert set(sd) == set(shapes)
        assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in sd)
    full = M.flux_dev_fill_params()
    n = sum((fi * fo_ + fo_) for _, fi, fo_ in M.linear_table(full))
    assert abs(n / 1e9 - 11.90) < 0.02            # 11.90 B base parameters


def test_schedule_api_and_packing_match_reference_goldens():
    from oracle import sampler_oracle as so
    from visualcloze_b200 import sampling, transport
    g = _golden("sampler.pt")
    for (n, L), ref in ((k, v) for k, v in g["get_schedule"].items() if isinstance(k, tuple)):
        assert sampling.get_schedule(n, L) == ref
        assert torch.equal(transport.solver_grid(0, 1, n + 1, L, True, 1), so.solver_grid(n + 1, L, True, 1))
    assert sampling.get_schedule(10, 1024, shift=False) == g["get_schedule"]["noshift"]
    assert torch.equal(sampling.time_shift(1.0416667, 1.0, torch.linspace(1, 0, 7)), g["time_shift"])
    assert [sampling.get_lin_function()(v) for v in (256, 3456, 4096)] == g["lin_fn"]
    # 30 time points at Li=3456 -> FLUX timesteps 1.0, 0.987554, ...
    tau = transport.solver_grid(0, 1, 30, 3456, True, 1)
    assert torch.allclose(1 - tau[:4], torch.tensor([1.0, 0.987554, 0.974528, 0.960878]), atol=2e-6)
    # SDEdit grid: strength 0.4, no shift
    sdedit = transport.Sampler(transport.create_transport()).sample_ode  # noqa: F841
    assert torch.allclose(transport.solver_grid(0.4, 1, 5, 64, False, 1.0), torch.linspace(0.4, 1, 5))
    pm = g["prepare_modified"]
    t5 = lambda prompts: torch.arange(len(prompts) * 6 * 8, dtype=torch.float32).reshape(len(prompts), 6, 8)
    clip = lambda prompts: torch.ones(len(prompts), 5)
    out = sampling.prepare_modified(t5, clip, pm["rows"], ["a", "b"], proportion_empty_prompts=0.0)
    assert set(out) == set(pm["out"])
    for k in pm["out"]:
        assert torch.equal(out[k], pm["out"][k]), k
    assert torch.equal(sampling.unpack(g["unpack"]["x"], 32, 96), g["unpack"]["out"])


def test_sampler_rejects_out_of_scope_and_cpu_inputs():
    from visualcloze_b200 import _lib, transport
    with pytest.raises(NotImplementedError):
        transport.create_transport("VP", "noise")
    s = transport.Sampler(transport.create_transport("Linear", "velocity"))
    with pytest.raises(NotImplementedError):
        s.sample_ode(sampling_method="dopri5")
    fn = s.sample_ode(sampling_method="euler", num_steps=4)
    with pytest.raises(_lib.VcbError):
        fn(torch.zeros(1, 8, 64, dtype=torch.bfloat16), lambda *a, **k: None, {})


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["REPO"])
from visualcloze_b200.parallel import gather_tiles, shard_samples
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{os.environ['PORT']}", rank=int(os.environ["RANK"]), world_size=2)
rank, n = dist.get_rank(), 5
def tile(s):
    g = torch.Generator().manual_seed(100 + s)
    return torch.randint(0, 255, (3, 4 + s, 6 + 2 * s), generator=g, dtype=torch.uint8)
mine = [tile(s) for s in shard_samples(n, rank, 2)]
allt = gather_tiles(mine, n)
assert len(allt) == n and all(torch.equal(allt[s], tile(s)) for s in range(n)), "gathered tiles differ from single-process tiles"
dist.barrier(); dist.destroy_process_group(); print("ok", rank)
"""


def test_gather_tiles_world2_gloo(tmp_path):
    """sample results are identical and ordered regardless of which rank computed them."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(os.environ, RANK=str(r), PORT=str(port), REPO=REPO),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    from visualcloze_b200.parallel import gather_tiles, shard_samples
    assert shard_samples(5, 1, 2) == [1, 3] and gather_tiles([torch.zeros(1, 2, 2)], 1)[0].shape == (1, 2, 2)


_SP_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["REPO"])
from visualcloze_b200.parallel import sp_gather_rows, sp_row_slice, sp_shard_rows
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{os.environ['PORT']}", rank=int(os.environ["RANK"]), world_size=2)
rank = dist.get_rank()
g = torch.Generator().manual_seed(3)
x = torch.randn(4, 1, 48, 8, generator=g)                 # [steps, B, Li, C] trajectory, identical on both ranks
loc = sp_shard_rows(x, rank, 2, dim=2)
assert loc.shape == (4, 1, 24, 8) and torch.equal(loc, x[:, :, 24 * rank: 24 * (rank + 1)])
row_local = loc * 2 + 1                                    # any row-local computation (the Euler update is one)
assert torch.equal(sp_gather_rows(row_local, dim=2), x * 2 + 1), "gathered rows differ from the unsharded computation"
assert sp_row_slice(512, rank, 2) == slice(256 * rank, 256 * (rank + 1))
try:
    sp_row_slice(49, rank, 2); raise SystemExit("indivisible row count must be rejected")
except ValueError:
    pass
dist.barrier(); dist.destroy_process_group(); print("ok", rank)
"""


def test_sequence_parallel_row_sharding_world2_gloo(tmp_path):
    """host side of the single-image sequence-parallel mode: row shards of both streams, gather = inverse."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "w_sp.py"
    script.write_text(_SP_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(os.environ, RANK=str(r), PORT=str(port), REPO=REPO),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    from visualcloze_b200 import parallel
    with pytest.raises(RuntimeError, match="process group"):
        parallel.SequenceParallel()


def test_pipeline_host_helpers_match_einops_and_reference_rules():
    from einops import rearrange
    from PIL import Image
    from visualcloze_b200 import pipeline as P
    m = torch.arange(2 * 32 * 48, dtype=torch.float32).reshape(2, 1, 32, 48)
    ref = rearrange(rearrange(m, "b c (h ph) (w pw) -> b (c ph pw) h w", ph=8, pw=8), "b c (h ph) (w pw) -> b (h w) (c ph pw)", ph=2, pw=2)
    assert torch.equal(P._pack_mask(m), ref)
    lat = torch.arange(16 * 4 * 6, dtype=torch.float32).reshape(1, 16, 4, 6)
    assert torch.equal(P._patchify(lat), rearrange(lat, "b c (h ph) (w pw) -> b (h w) (c ph pw)", ph=2, pw=2))
    # area ~ res^2, sides multiples of 16, square stays res x res
    assert P.resize_with_aspect_ratio(Image.new("RGB", (500, 500)), 384).size == (384, 384)
    w, h = P.resize_with_aspect_ratio(Image.new("RGB", (800, 400)), 384).size
    assert w % 16 == 0 and h % 16 == 0 and abs(w * h - 384 * 384) / (384 * 384) < 0.12
    rgba = Image.new("RGBA", (4, 4), (10, 20, 30, 0))
    assert P.to_rgb_if_rgba(rgba).getpixel((0, 0)) == (255, 255, 255)
    t = P.image_transform(Image.new("RGB", (4, 2), (255, 0, 127)))
    assert t.shape == (3, 2, 4) and float(t[0].max()) == 1.0 and float(t[1].min()) == -1.0
    pm = object.__new__(P.VisualClozeModel)
    pm.grid_h, pm.grid_w, pm.resolution = 2, 3, 64
    imgs = [[Image.new("RGB", (100, 100), (i, j, 0)) for j in range(3)] for i in range(2)]
    imgs[1][2] = None
    proc, mask_pos, up = pm._prepare_grid(imgs)
    assert len(proc) == 6 and mask_pos == [0, 0, 1] and up == (100, 100) and all(p.size == (64, 64) for p in proc)
    imgs[0][1] = None
    with pytest.raises(ValueError, match="in-context"):
        pm._prepare_grid(imgs)


def test_never_loaded_weights_are_refused():
    """a drop-in constructor must not leave torch.empty parameters behind.  The engines refuse to pack parameters that
    never received values; load_state_dict / init_synthetic clear the guard; LoRA and QK-norm tensors keep their defined defaults."""
    from visualcloze_b200 import _lib, model as M, vae as V
    small = dict(in_channels=384, out_channels=64, vec_in_dim=32, context_in_dim=64, hidden_size=256, mlp_ratio=2.0, num_heads=2,
                 depth=1, depth_single_blocks=1, axes_dim=[16, 56, 56], theta=10_000, qkv_bias=True, guidance_embed=True)
    m = M.FluxLoraWrapper(lora_rank=8, params=M.FluxParams(**small))
    assert len(m.uninitialized()) > 0 and all(".lora_" not in n and not n.endswith(".scale") for n in m.uninitialized())
    with pytest.raises(_lib.VcbError, match="never loaded"):
        m.engine()
    assert float(m.state_dict()["img_in.lora_A.weight"].abs().max()) == 0.0
    base = {k: torch.zeros_like(v) for k, v in m.state_dict().items() if ".lora_" not in k}
    m.load_state_dict(base, strict=False)
    assert m.uninitialized() == []
    d = V.AutoEncoderDecoder(V.AutoEncoderParams(ch=64, ch_mult=[1, 2], num_res_blocks=1))
    with pytest.raises(_lib.VcbError, match="never loaded"):
        d._engine()


def test_linear_precision_levels_and_the_header_constants_agree():
    """set_linear_precision takes "bf16" | "fp8" | "fp8_all" and nothing else; the levels the Python side passes to
    vcb_flux_set_fp8 are the VCB_FP8_* constants of include/vcb200.h; the row quantiser is part of the exported ABI."""
    import re
    from visualcloze_b200 import _lib, model as M
    small = dict(in_channels=384, out_channels=64, vec_in_dim=32, context_in_dim=64, hidden_size=256, mlp_ratio=2.0, num_heads=2,
                 depth=1, depth_single_blocks=1, axes_dim=[16, 56, 56], theta=10_000, qkv_bias=True, guidance_embed=True)
    m = M.FluxLoraWrapper(lora_rank=8, params=M.FluxParams(**small))
    assert m.linear_precision == "bf16"
    for p in ("fp8", "fp8_all", "bf16"):
        m.set_linear_precision(p)
        assert m.linear_precision == p
    for bad in ("fp16", "FP8", "", None):
        with pytest.raises(ValueError):
            m.set_linear_precision(bad)
    hdr = open(os.path.join(REPO, "include", "vcb200.h")).read()
    consts = {k: int(v) for k, v in re.findall(r"#define (VCB_FP8_[A-Z_]+) (\d+)", hdr)}
    assert consts == {"VCB_FP8_OFF": 0, "VCB_FP8_LN_FED": 1, "VCB_FP8_ALL_LINEARS": 2}
    src = open(os.path.join(REPO, "visualcloze_b200", "model.py")).read()
    assert '{"bf16": 0, "fp8": 1, "fp8_all": 2}' in src
    assert hasattr(_lib.lib(), "vcb_quantize_rows_e4m3")


def test_text_encoder_host_logic():
    """the T5 relative-position bucket rule equals HF's (modeling_t5.T5Attention._relative_position_bucket), the
    encoders refuse a CPU device (no fallback), and the reference-shaped HFEmbedder wrapper returns what conditioner.py:10 selects."""
    from transformers.models.t5.modeling_t5 import T5Attention
    from visualcloze_b200 import _lib, text_encoders as T
    rp = torch.arange(-700, 700)[None, :] - torch.arange(0, 3)[:, None]
    for nb, md in ((32, 128), (16, 64)):
        assert torch.equal(T.t5_relative_position_bucket(rp, nb, md),
                           T5Attention._relative_position_bucket(rp, bidirectional=True, num_buckets=nb, max_distance=md))
    with pytest.raises(_lib.VcbError, match="no CPU fallback"):
        T.T5Encoder({}, num_heads=4, num_layers=0, device="cpu")
    with pytest.raises(ValueError):
        T.T5Encoder({}, num_heads=4, num_layers=0, d_kv=128)

    class Tok:                                   # tokenizer stand-in: records the reference's call (conditioner.py:23-31)
        def __call__(self, text, **kw):
            self.kw = kw
            return {"input_ids": torch.zeros(len(text), kw["max_length"], dtype=torch.long)}
    enc = T.CLIPTextEncoder.__new__(T.CLIPTextEncoder)
    enc.__call__ = None
    tok = Tok()
    emb = T.HFEmbedder(tok, lambda ids: ("last", "pooled"), 77)
    emb.is_clip = True
    assert emb(["a", "b"]) == "pooled" and tok.kw["padding"] == "max_length" and tok.kw["truncation"] is True and tok.kw["max_length"] == 77
    emb.is_clip = False
    emb.encoder = lambda ids: "hidden"
    assert emb.forward(["a"]) == "hidden"


def test_diffusers_vae_keys_convert_to_the_reference_names():
    """visualcloze.py:100 loads diffusers' AutoencoderKL; our VAE modules carry the in-repo AutoEncoder names (autoencoder.py).
    The converter must map a diffusers-named state dict of the FLUX VAE geometry onto exactly our parameter set."""
    from visualcloze_b200 import vae as V
    p = V.AutoEncoderParams()
    ours = {**V.decoder_param_shapes(p), **V.encoder_param_shapes(p)}

    def to_diffusers(k):
        half, rest = k.split(".", 1)
        rest = re.sub(r"^mid\.block_(\d)\.", lambda m: f"mid_block.resnets.{int(m.group(1)) - 1}.", rest)
        rest = re.sub(r"^mid\.attn_1\.", "mid_block.attentions.0.", rest)
        for a, b in (("attentions.0.norm.", "attentions.0.group_norm."), ("attentions.0.q.", "attentions.0.to_q."), ("attentions.0.k.", "attentions.0.to_k."),
                     ("attentions.0.v.", "attentions.0.to_v."), ("attentions.0.proj_out.", "attentions.0.to_out.0.")):
            rest = rest.replace(a, b)
        rest = re.sub(r"^down\.(\d+)\.block\.(\d+)\.", r"down_blocks.\1.resnets.\2.", rest)
        rest = re.sub(r"^down\.(\d+)\.downsample\.conv\.", r"down_blocks.\1.downsamplers.0.conv.", rest)
        rest = re.sub(r"^up\.(\d+)\.block\.(\d+)\.", lambda m: f"up_blocks.{3 - int(m.group(1))}.resnets.{m.group(2)}.", rest)
        rest = re.sub(r"^up\.(\d+)\.upsample\.conv\.", lambda m: f"up_blocks.{3 - int(m.group(1))}.upsamplers.0.conv.", rest)
        rest = rest.replace("nin_shortcut", "conv_shortcut").replace("norm_out.", "conv_norm_out.")
        return f"{half}.{rest}"

    g = torch.Generator().manual_seed(0)
    src, expect = {}, {}
    for k, shp in ours.items():
        t = torch.randn(shp, generator=g)
        dk = to_diffusers(k)
        expect[k] = t
        # diffusers stores the attention projections as Linear [C, C]
        src[dk] = t[:, :, 0, 0] if ("attentions.0.to_" in dk and dk.endswith("weight")) else t
    src["quant_conv.weight"] = torch.zeros(1)
    conv = V.diffusers_vae_to_bfl(src)
    assert set(conv) == set(ours)
    assert all(torch.equal(conv[k], expect[k]) for k in ours)
    d = V.AutoEncoderDecoder()
    d.load_state_dict(src, strict=True)               # diffusers names, both halves present: decoder half taken
    assert d.uninitialized() == []


def test_engine_rejects_masks_the_kernel_cannot_honour():
    """a padded txt_mask or a non-prefix img_mask must raise, not silently attend to padded rows.  (CPU check of
    the validation rule itself; the GPU tests cover the accepted right-padded layout.)"""
    import inspect
    from visualcloze_b200 import engine
    src = inspect.getsource(engine.FluxEngine.prepare)
    assert "unsupported attention mask" in src and "prefix" in src


def test_attention_persistent_schedule_covers_every_key_tile_exactly_once():
    """attn4_sm100.cuh's schedule (AttnSched / AttnSegIter), restated: full rounds dealt out like the per-pair grid, the partial last
    round cut along the key tiles into equal shares.  The shares must tile the (unit, key tile) space exactly once; a piece that
    does not hold key tile 0 must be its CTA's FIRST phase-2 segment (it waits on nothing), a cut unit's finaliser must be its
    CTA's LAST segment, and its contributors must be the next CTAs, contiguous up to the last key tile -- the deadlock-freedom
    argument of the kernel rests on exactly these properties; the host-side bound on the segment list must hold."""
    def sched(n_heads, n_qt, sms=148):
        n_pairs, n_kv = (n_qt + 1) // 2, n_qt
        n_units = n_heads * n_pairs
        G = sms if n_units * n_kv // sms >= 4 else max(n_units * n_kv // 4, 1)
        R, rem = divmod(n_units, G)
        U = rem * n_kv
        G2 = 0 if rem == 0 else min(G, max(U // 4, 1))
        return n_pairs, n_kv, n_units, G, R, rem, U, G2

    def boundary(c, S):
        n_pairs, n_kv, n_units, G, R, rem, U, G2 = S
        ur, kv = divmod(U * c // G2, n_kv)
        return R * G + ur, kv

    def segs(c, S):
        n_pairs, n_kv, n_units, G, R, rem, U, G2 = S
        out = [(r * G + c, 0, n_kv, 1) for r in range(R)]
        if c < G2:
            (us, ks), (ue, ke) = boundary(c, S), boundary(c + 1, S)
            for u in range(us, min(ue, n_units - 1) + 1):
                kv0, kv1 = (ks if u == us else 0), (ke if u == ue else n_kv)
                if kv0 < kv1:
                    out.append((u, kv0, kv1, 2))
        return out

    for B, H, L in [(1, 24, 3968), (1, 24, 7424), (1, 24, 4608), (1, 24, 6656), (1, 24, 1088), (1, 2, 128), (1, 2, 200), (1, 2, 1088),
                    (1, 3, 1500), (2, 2, 520), (1, 1, 640), (8, 24, 1088), (4, 24, 256), (1, 3, 3968), (1, 4, 2000), (8, 24, 3968),
                    (1, 37, 640), (1, 149, 256)]:
        n_qt = (L + 127) // 128
        S = sched(B * H, n_qt)
        n_pairs, n_kv, n_units, G, R, rem, U, G2 = S
        cover = set()
        for c in range(G):
            sg = segs(c, S)
            assert len(sg) <= n_units // G + 4 <= 64 or n_units // G + 4 > 64, "host-side bound on the segment list"
            assert len(sg) <= n_units // G + 4
            tail = [x for x in sg if x[3] == 2]
            for i, (u, a, b, ph) in enumerate(tail):
                assert b - a >= 1
                assert a == 0 or i == 0
                if a == 0 and b < n_kv:
                    assert i == len(tail) - 1
                    pos, pc = b, c + 1
                    while pc < G2 and boundary(pc, S)[0] == u:
                        f = [x for x in segs(pc, S) if x[3] == 2][0]
                        assert f[0] == u and f[1] == pos and f[1] > 0
                        pos, pc = f[2], pc + 1
                    assert pos == n_kv
            for (u, a, b, ph) in sg:
                for k in range(a, b):
                    assert (u, k) not in cover
                    cover.add((u, k))
        assert len(cover) == n_units * n_kv, (B, H, L, len(cover), n_units * n_kv)


def test_diffusers_style_adapter_maps_the_documented_call():
    """README.md:141-205 of the reference documents the diffusers ``VisualClozePipeline`` call; the adapter must turn it into the
    reference entry's arguments (grid size, three prompts, seed from the generator, SDEdit switch) and wrap the result like
    ``.images[0][0]``."""
    from PIL import Image
    from visualcloze_b200.diffusers_adapter import VisualClozePipelineAdapter, layout_prompt

    class Fake:
        max_length = 512

        def set_grid_size(self, h, w):
            self.grid = (h, w)

        def process_images(self, images, prompts, **kw):
            self.call = dict(images=images, prompts=prompts, **kw)
            return [Image.new("RGB", (64, 64))]

    im = Image.new("RGB", (32, 32))
    fake = Fake()
    pipe = VisualClozePipelineAdapter(fake)
    out = pipe(task_prompt="do the task", content_prompt=None, image=[[im, im, im], [im, im, None]], upsampling_height=160,
               upsampling_width=128, upsampling_strength=0.3, guidance_scale=30, num_inference_steps=30, max_sequence_length=512,
               generator=torch.Generator("cpu").manual_seed(7))
    assert fake.grid == (2, 3) and fake.call["seed"] == 7 and fake.call["cfg"] == 30 and fake.call["steps"] == 30
    assert fake.call["prompts"] == [layout_prompt(2, 3), "do the task", ""] and fake.call["is_upsampling"] is True
    assert abs(fake.call["upsampling_noise"] - 0.3) < 1e-9 and fake.call["images"][1][2] is None
    assert out.images[0][0].size == (128, 160)
    assert layout_prompt(2, 3) == "A grid layout with 2 rows and 3 columns, displaying 6 images arranged side by side."
    pipe(task_prompt="t", content_prompt="c", image=[[im, None]], upsampling_strength=1.0)
    assert fake.call["is_upsampling"] is False and fake.grid == (1, 2)
    with pytest.raises(ValueError, match="in-context"):
        pipe(task_prompt="t", content_prompt="c", image=[[im, None], [im, None]])
