"""Condense round-2 evidence (gpurun_out/ + profiles/r02_*.json) into profiles/r02_summary.md and profiles/r02_ncu_traffic.json.
Runs in the build container (reads the .ncu-rep files with `ncu -i ... --page raw --csv`).  Usage: python tools/make_r02_summary.py"""
import collections
import csv
import glob
import io
import json
import os
import subprocess

G, P = "gpurun_out", "profiles"
out = ["# r02 — measured evidence (B200, sm_100a)\n",
       "\nEvery number below comes from a `gpurun` session of this round; ncu numbers are never bench values. The chip runs the "
       "denoising loop at its ~1 kW software power cap (SM clock 1425-1470 MHz of 1965), so `kernel_ms x MHz` is the box-independent quantity.\n"]


def section(t):
    out.append(f"\n## {t}\n\n")


def jl(path):
    return json.loads(open(path).read().strip().splitlines()[-1]) if os.path.exists(path) else None


# ---- bench lines -------------------------------------------------------------------------------------------------
rows = []
for tag, path in (("bf16 (headline)", f"{P}/r02_bench_1gpu.json"), ("fp8 LayerNorm-fed Linears (opt-in)", f"{P}/r02_bench_fp8_1gpu.json"),
                  ("fp8 every block Linear (opt-in, fp8_all)", f"{P}/r02_bench_fp8_all_1gpu.json"),
                  ("bf16, VCB_STREAMK=1", f"{P}/r02_bench_streamk_1gpu.json")):
    if os.path.exists(path):
        rows.append((tag, json.load(open(path))))
if rows:
    section("bench.py, 1 GPU, cfg B (3 timed images)")
    out.append("| line | images/s | e2e | ms/image | SM MHz | GEMM ms / frac | attention ms / frac | LN ms / frac | VAE ms | exact-softmax images/s |\n|---|---|---|---|---|---|---|---|---|---|\n")
    for tag, b in rows:
        k = b["kernel_time_share"]
        out.append(f"| {tag} | {b['value']:.4f} | {b['e2e']['value']:.4f} | {b['ms_per_step']:.0f} | {b['clocks']['sm_mhz']:.0f} | "
                   f"{k['gemm_ms']:.0f} / {b['roofline']['frac']:.3f} | {k['attention_ms']:.0f} / {b['roofline_attention']['frac']:.3f} | "
                   f"{k['ln_modulate_ms']:.1f} / {b['roofline_ln_modulate']['frac']:.3f} | {b['roofline_vae']['decode_ms_total']:.1f} | "
                   f"{b['extra']['attention_exact_softmax']['images_per_s']:.4f} |\n")
    b = rows[0][1]
    out.append(f"\nsoftmax variant per block: `{json.dumps(b['roofline_attention']['softmax_variant_per_block'])}`; `roofline_vae`: `{json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in b['roofline_vae'].items() if k != 'kernel'})}`\n")
    out.append("\nPer-shape GEMM table of the headline line (CUDA events around every launch of one instrumented image):\n\n| M | N | K | epilogue | tile | launches | avg us | TFLOP/s | frac of 1453 | share |\n|---|---|---|---|---|---|---|---|---|---|\n")
    for r in b["gemm_shapes"]:
        out.append(f"| {r['M']} | {r['N']} | {r['K']} | {r['epilogue']} | {r['block_n']}x{r['cta_group']}cta | {r['launches']} | {r['avg_us']:.1f} | {r['tflops']:.0f} | "
                   f"{r['frac_of_sustained_peak']:.3f} | {100 * r['share_of_gemm_time']:.1f}% |\n")

for name, title in (("r02_bench_n2", "bench.py --gpus 2 (replicas + extra.sp)"), ("r02_bench_n2_sp_persist", "same, VCB_SP_ATTN_PERSIST=1"),
                    ("r02_bench_B_n8", "bench.py --gpus 8, cfg B"), ("r02_bench_C_n8", "bench.py --gpus 8 --workload C (BASELINE config 3)"),
                    ("r02_bench_E_n4", "bench.py --gpus 4 --workload E (BASELINE config 5)")):
    path = f"{P}/{name}.json"
    if os.path.exists(path):
        b = json.load(open(path))
        section(title)
        out.append(f"value {b['value']:.4f} images/s over {b['n_gpus']} GPUs ({b['ms_per_step']:.0f} ms per step, {b['steps']} timed), e2e {b['e2e']['value']:.4f}, "
                   f"SM {b['clocks']['sm_mhz']} MHz; workload `{b['config']['workload']}`\n\n`extra.sp` = `{json.dumps(b['extra'].get('sp'))}`\n")

# ---- library comparisons -------------------------------------------------------------------------------------------
p = f"{P}/r02_gemm_shapes_sustained_vs_cublas.json"
if os.path.exists(p):
    d = json.load(open(p))
    section("the six cfg-B GEMM shapes, ours WITH fused epilogues vs cuBLAS WITHOUT, each back to back for 0.6 s (power-capped steady state)")
    out.append("| shape | M | N | K | ours us | ours TFLOP/s | cuBLAS us | cuBLAS TFLOP/s | cuBLAS time / ours |\n|---|---|---|---|---|---|---|---|---|\n")
    for r in d["rows"]:
        out.append(f"| {r['name']} | {r['M']} | {r['N']} | {r['K']} | {r['ours_us']:.1f} | {r['ours_tflops']:.0f} | {r['cublas_us']:.1f} | {r['cublas_tflops']:.0f} | {r['ours_over_cublas']:.3f} |\n")
    out.append(f"\nper evaluation: ours {d['per_eval_ms']['ours']:.2f} ms, cuBLAS without epilogues {d['per_eval_ms']['cublas_without_epilogues']:.2f} ms\n")
p = f"{P}/r02_attn_vs_libs.json"
if os.path.exists(p):
    d = json.load(open(p))
    section("attention, 24 heads, head_dim 128: TFLOP/s alone (L2 flushed, burst clocks) / sustained (back to back 0.5 s)")
    keys = ["vcb_per_pair_bounded", "vcb_per_pair_exact", "vcb_persistent_bounded", "vcb_persistent_exact", "cudnn_sdpa", "flash_attn2"]
    out.append("| L | " + " | ".join(keys) + " |\n|---|" + "---|" * len(keys) + "\n")
    for r in d["rows"]:
        cells = []
        for k in keys:
            v = r.get(k, {})
            cells.append(f"{v['tflops']:.0f} / {v.get('sustained_tflops', float('nan')):.0f}" if "tflops" in v else "-")
        out.append(f"| {r['L']} | " + " | ".join(cells) + " |\n")
    out.append(f"\nmethod: {d['method']}\n")

p = f"{P}/r02_text_encoders.json"
if os.path.exists(p):
    d = json.load(open(p))
    section("text encoders, once per image (tools/bench_text_encoders.py; real geometry, random weights; HF bf16 is a reference point only)")
    for k, v in d.items():
        out.append(f"* `{k}`: " + ", ".join(f"{a} = {b:.4g}" if isinstance(b, float) else f"{a} = {b}" for a, b in v.items()) + "\n")

# ---- parity --------------------------------------------------------------------------------------------------------
for name, title in (("r02_fullsize_parity", "full-size parity (tests/test_fullsize_gpu.py)"), ("r02_fp8_parity", "fp8 vs bf16 projections (tests/test_fp8_gpu.py)")):
    p = f"{P}/{name}.json"
    if os.path.exists(p):
        section(title)
        for k, v in json.load(open(p)).items():
            out.append(f"* `{k}`: " + ", ".join(f"{a} = {b:.4g}" for a, b in v.items()) + "\n")

# ---- ncu launch list -----------------------------------------------------------------------------------------------
LL = f"{P}/r02_ncu_launches.csv" if os.path.exists(f"{P}/r02_ncu_launches.csv") else f"{G}/r2_launches.csv"
if os.path.exists(LL):
    rows = list(csv.reader(open(LL)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"]
    if hi:
        hdr, data = rows[hi[0]], rows[hi[0] + 1:]
        ix = {n: i for i, n in enumerate(hdr)}
        agg = collections.OrderedDict()
        for r in data:
            if len(r) < len(hdr):
                continue
            v = float(r[ix["Metric Value"]].replace(",", ""))
            u = r[ix["Metric Unit"]]
            v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
            k = (r[ix["Kernel Name"]].split("(")[0][:80], r[ix["Grid Size"]])
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += v
        tot = sum(v[1] for v in agg.values())
        section("ncu launch list: `ncu --metrics gpu__time_duration.sum --clock-control none` over `tools/time_full.py 3` (prepare + 2 evaluations, cfg B; cold-cache, serialised: compare SHARES)")
        out.append("| share | total us | launches | avg us | kernel | grid |\n|---|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:20]:
            out.append(f"| {100 * v[1] / tot:.1f}% | {v[1]:.0f} | {v[0]} | {v[1] / v[0]:.1f} | `{k[0]}` | {k[1]} |\n")
        out.append(f"\ntotal {tot / 1e3:.1f} ms over {sum(v[0] for v in agg.values())} launches\n")

# ---- ncu --set full captures ---------------------------------------------------------------------------------------
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "ns": 1e-3, "us": 1, "ms": 1e3}
traffic = {}
caps = sorted(glob.glob(f"{G}/r2_prof_*.raw.csv")) or sorted(glob.glob(f"{P}/r02_ncu_*.raw.csv"))
if caps:
    section("`ncu --set full --clock-control none --import-source on`, one launch per hot kernel at its cfg-B shape (`tools/ncu_targets.py`; cold, single launch)")
    out.append("| capture | kernel | grid | time us | DRAM read+write MB | tensor pipe % | DRAM % | SM % | regs | L2 hit % | warps active % | XU % | issue active % |\n|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for p in caps:
        name = os.path.basename(p).replace("r2_prof_", "").replace("r02_ncu_", "").replace(".raw.csv", "")
        rows = list(csv.reader(open(p)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]

        def val(r, n):
            if n not in hdr:
                return float("nan")
            i = hdr.index(n)
            try:
                return float(r[i].replace(",", "")) * mult.get(units[i], 1)
            except ValueError:
                return float("nan")
        for r in rows[2:]:
            dram = (val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"))
            t = val(r, "gpu__time_duration.sum")
            out.append(f"| {name} | `{r[hdr.index('Kernel Name')][:48]}` | {r[hdr.index('Grid Size')]} | {t:.1f} | {dram / 1e6:.1f} | "
                       + " | ".join(f"{val(r, n):.1f}" for n in WANT[3:]) + " |\n")
            key = {"gemm_linear2": "gemm", "attn_pair": "attention", "ln_stats": "ln_modulate", "conv256": "vae_conv3x3"}.get(name)
            if key:
                traffic[key] = {"avg_dram_bytes_per_launch": dram, "time_us": t, "tensor_pipe_active_pct": val(r, WANT[3]), "capture": os.path.basename(p)}
if traffic:
    traffic["note"] = "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from `ncu --set full --clock-control none` (profiles/r02_summary.md); cold caches"
    json.dump(traffic, open(f"{P}/r02_ncu_traffic.json", "w"), indent=1)

open(f"{P}/r02_summary.md", "w").write("".join(out))
print(f"wrote {P}/r02_summary.md ({sum(len(s) for s in out)} bytes)")
