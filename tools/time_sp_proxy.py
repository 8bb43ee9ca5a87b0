"""Single-GPU proxy for one rank of the sequence-parallel mode at world W: the denoise loop on Li/W + Lt/W tokens has exactly
that rank's GEMM / LayerNorm shapes (attention differs: all heads over L/W rows instead of heads/W over L rows).  Prints the
per-category kernel times of one image so GEMM tile / stream-K policies can be compared in the loop (VCB_STREAMK=0|1|unset)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visualcloze_b200 import _lib, model as M, transport as T  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
BF16 = torch.bfloat16
with torch.device("cuda"):
    model = M.FluxLoraWrapper(lora_rank=256, params=M.flux_dev_fill_params())
model.init_synthetic(0)
model.engine()
Li, Lt = 3456 // W, 512 // W
g = torch.Generator().manual_seed(1234)
ids = torch.zeros(1, Li, 3)
ids[0, :, 0] = 1
ids[0, :, 1] = torch.arange(Li) // 72
ids[0, :, 2] = torch.arange(Li) % 72
x = torch.randn(1, Li, 64, generator=g).to(BF16).cuda()
kw = dict(txt=(0.1 * torch.randn(1, Lt, 4096, generator=g)).to(BF16).cuda(), txt_ids=torch.zeros(1, Lt, 3).cuda(),
          txt_mask=torch.ones(1, Lt, dtype=torch.int32).cuda(), y=torch.randn(1, 768, generator=g).to(BF16).cuda(),
          img_ids=ids.cuda(), img_mask=torch.ones(1, Li, dtype=torch.int32).cuda(),
          cond=torch.randn(1, Li, 320, generator=g).to(BF16).cuda(), guidance=torch.full((1,), 30.0, dtype=BF16).cuda())
fn = T.Sampler(T.create_transport("Linear", "velocity", do_shift=True)).sample_ode(
    sampling_method="euler", num_steps=steps, atol=1e-6, rtol=1e-3, reverse=False, do_shift=True, time_shifting_factor=1)
lib = _lib.lib()
for it in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    traj = fn(x, model.forward, kw)
    e1.record()
    torch.cuda.synchronize()
    print(f"iter {it}: {e0.elapsed_time(e1):.1f} ms/image-shard (W={W}, Li={Li}, Lt={Lt}), finite={bool(torch.isfinite(traj[-1].float()).all())}", flush=True)
lib.vcb_profile_begin()
fn(x, model.forward, kw)
ms, n = (C.c_double * 4)(), (C.c_longlong * 4)()
lib.vcb_profile_end(ms, n)
print(f"VCB_STREAMK={os.environ.get('VCB_STREAMK', 'auto')}: gemm {ms[0]:.1f} ms ({n[0]}), attention {ms[1]:.1f} ms, ln {ms[2]:.1f} ms, other {ms[3]:.1f} ms", flush=True)
