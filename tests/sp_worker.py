"""Rank program of tests/test_sp_gpu.py (run under torch.distributed.run, one process per GPU, NCCL).

Checks the single-image sequence-parallel mode against the single-GPU path of the same library and
against the CPU oracle / reference goldens; rank 0 writes a JSON report to argv[1].
"""
import ctypes as C
import json
import os
import sys
import time

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
BF16 = torch.bfloat16


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def main():
    out_path = sys.argv[1]
    full = len(sys.argv) > 2 and sys.argv[2] == "full"
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    from visualcloze_b200 import _lib, parallel
    import visualcloze_b200.model as m
    import visualcloze_b200.transport as t
    from oracle import flux_oracle as fo
    lib = _lib.lib()
    st = lambda: torch.cuda.current_stream().cuda_stream
    rep = {"world": world}

    # ---- 1. peer-mapped buffers, NVLink stores, phase barrier -----------------------------------------
    rows, cols = world, 64
    pb = parallel.PeerBuffer(rows * cols * 2)
    fl = parallel.PeerBuffer(256)
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    src = torch.full((1, cols), float(rank + 1), dtype=BF16, device="cuda")
    for r in range(world):      # my row -> row `rank` of every rank's buffer (peer store from an ordinary kernel)
        _lib.check(lib.vcb_copy_cols(src.data_ptr(), cols, pb.ptrs[r] + rank * cols * 2, cols, 0, 1, cols, st()), "copy_cols")
    _lib.check(lib.vcb_sp_barrier(fl.array(), world, rank, 1, err.data_ptr(), 5000, st()), "sp_barrier")
    got = torch.empty(rows, cols, dtype=BF16, device="cuda")
    _lib.check(lib.vcb_copy_cols(pb.local, cols, got.data_ptr(), cols, 0, rows, cols, st()), "copy_cols")
    torch.cuda.synchronize()
    want = torch.arange(1, world + 1, dtype=torch.float32)[:, None].expand(rows, cols)
    rep["peer_store_ok"] = bool(torch.equal(got.float().cpu(), want)) and int(err.item()) == 0
    # a second epoch, and the timeout path: a barrier nobody else joins must report instead of hanging
    _lib.check(lib.vcb_sp_barrier(fl.array(), world, rank, 2, err.data_ptr(), 5000, st()), "sp_barrier")
    torch.cuda.synchronize()
    rep["barrier2_ok"] = int(err.item()) == 0
    if rank == 0:
        t0 = time.time()
        _lib.check(lib.vcb_sp_barrier(fl.array(), world, rank, 1000, err.data_ptr(), 200, st()), "sp_barrier")
        torch.cuda.synchronize()
        rep["timeout_reported"] = int(err.item()) == 1000 and (time.time() - t0) < 3.0
    dist.barrier()
    pb.close()
    fl.close()

    # ---- 2. small model: Flux.forward, SP vs single GPU vs oracle vs reference golden ---------------------
    from test_flux_gpu import _build, _cuda, _load
    g = _load("flux_small_b1.pt")
    cfg_dict = dict(g["cfg"])
    pinned = cfg_dict["num_heads"] % world == 0        # the reference-generated golden is for the 2-head geometry
    if not pinned:                                     # more ranks than heads: widen to one head per rank (oracle-checked only)
        cfg_dict.update(num_heads=world, hidden_size=128 * world)
    cfg, params, model = _build(m, cfg_dict, g["param_seed"])
    inp = g["inputs"]
    single = model(**_cuda(inp)).cpu()
    sp = parallel.SequenceParallel()
    model.enable_sequence_parallel(sp)
    spo = model(**_cuda(inp)).cpu()
    ref = fo.flux_forward(params, cfg, **inp, mode="cuda_bf16")
    rep["small_sp_vs_single"] = rel_l2(spo, single)
    rep["small_sp_vs_oracle"] = rel_l2(spo, ref)
    rep["small_sp_vs_golden"] = rel_l2(spo, g["out_cpu_bf16"]) if pinned else None
    allo = [torch.empty_like(spo).cuda() for _ in range(world)]
    dist.all_gather(allo, spo.cuda())
    rep["small_ranks_identical"] = all(torch.equal(a, allo[0]) for a in allo)

    # ---- 3. sampler trajectory through the public API (sharded Euler loop + final gather) ------------------
    gs = _load("sampler.pt")["shift4"]
    kw = dict(gs["kwargs"])
    sampler = t.Sampler(t.create_transport("Linear", "velocity", do_shift=True))
    fn = sampler.sample_ode(sampling_method="euler", atol=1e-6, rtol=1e-3, reverse=False, **kw)
    mk = dict(_cuda(gs["inputs"]), cond=gs["cond"].cuda())
    traj_sp = fn(gs["x"].cuda(), model.forward, mk).cpu()
    model.enable_sequence_parallel(None)
    traj_1 = fn(gs["x"].cuda(), model.forward, mk).cpu()
    rep["traj_shape_ok"] = tuple(traj_sp.shape) == tuple(gs["traj"].shape)
    rep["traj_sp_vs_single"] = rel_l2(traj_sp[-1], traj_1[-1])
    rep["traj_sp_vs_golden"] = rel_l2(traj_sp[-1], gs["traj"][-1]) if pinned else None
    rep["traj_x0_exact"] = bool(torch.equal(traj_sp[0], gs["x"]))

    # ---- 4. FLUX width (hidden 3072, 24 heads) at depth 1+1 on the cfg-B token count: parity + per-eval timing ------
    if full:
        from visualcloze_b200.model import FluxLoraWrapper, flux_dev_fill_params
        P = flux_dev_fill_params()
        P.depth, P.depth_single_blocks = 2, 4
        with torch.device("cuda"):
            big = FluxLoraWrapper(lora_rank=16, params=P).init_synthetic(3)
        gen = torch.Generator().manual_seed(11)
        gh, gw, res = 2, 3, 384
        h, w = res // 16, gw * res // 16
        ids = []
        for j in range(gh):
            tt = torch.zeros(h, w, 3)
            tt[..., 0] = j + 1
            tt[..., 1] += torch.arange(h)[:, None]
            tt[..., 2] += torch.arange(w)[None, :]
            ids.append(tt.reshape(-1, 3))
        ids = torch.cat(ids)[None]
        Li, Lt = ids.shape[1], 512
        binp = dict(img=torch.randn(1, Li, 384, generator=gen).to(BF16), img_ids=ids,
                    txt=(0.1 * torch.randn(1, Lt, 4096, generator=gen)).to(BF16), txt_ids=torch.zeros(1, Lt, 3),
                    timesteps=torch.tensor([0.63]), y=torch.randn(1, 768, generator=gen).to(BF16),
                    txt_mask=torch.ones(1, Lt, dtype=torch.int32), img_mask=torch.ones(1, Li, dtype=torch.int32),
                    guidance=torch.full((1,), 30.0, dtype=BF16))
        cin = _cuda(binp)
        one = big(**cin)
        big.enable_sequence_parallel(sp)
        two = big(**cin)
        rep["big_sp_vs_single"] = rel_l2(two.cpu(), one.cpu())

        def time_evals(n=6):
            eng = big.engine()
            big(**cin)                                            # prepare + warm
            img_l = eng._sp.shard(cin["img"]).contiguous() if eng._sp is not None else cin["img"]
            for _ in range(2):
                eng.forward(0, img_l)
            torch.cuda.synchronize()
            dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                eng.forward(0, img_l)
            b.record()
            torch.cuda.synchronize()
            ms = torch.tensor([a.elapsed_time(b) / n], device="cuda")
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms)

        rep["big_ms_per_eval_sp"] = time_evals()
        big.enable_sequence_parallel(None)
        rep["big_ms_per_eval_single"] = time_evals()
        rep["big_blocks"] = "2 double + 4 single, L=3968"
    sp.release()
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(rep, f, indent=1)
        print(json.dumps(rep))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
