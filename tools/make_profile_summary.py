"""Condense gpurun_out/ evidence (bench line, ncu launch list, ncu full-set reports, kernel micro-benchmarks) into
tracked markdown under profiles/.  Usage: python tools/make_profile_summary.py r01"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G = "gpurun_out"
out = [f"# {tag} — measured evidence (B200, sm_100a)\n"]


def section(t):
    out.append(f"\n## {t}\n")


if os.path.exists(f"{G}/bench_{tag}_n1.json"):
    b = json.loads(open(f"{G}/bench_{tag}_n1.json").read().strip().splitlines()[-1])
    section("bench.py line (N=1)")
    out.append("```json\n" + json.dumps(b, indent=1) + "\n```\n")

if os.path.exists(f"{G}/launches.csv"):
    rows = list(csv.reader(open(f"{G}/launches.csv")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ix = {n: i for i, n in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in data:
        if len(r) < len(hdr):
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        u = r[ix["Metric Unit"]]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        k = (r[ix["Kernel Name"]].split("(")[0][:70], r[ix["Grid Size"]])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    section("ncu launch list: `ncu --metrics gpu__time_duration.sum --clock-control none` over `tools/time_full.py 3` "
            "(prepare + 2 model evaluations, cfg B). Cold-cache, serialised: compare SHARES")
    out.append("| share | total us | launches | avg us | kernel | grid |\n|---|---|---|---|---|---|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
        out.append(f"| {100 * v[1] / tot:.1f}% | {v[1]:.0f} | {v[0]} | {v[1] / v[0]:.1f} | `{k[0]}` | {k[1]} |\n")
    out.append(f"\ntotal {tot / 1e3:.1f} ms over {sum(v[0] for v in agg.values())} launches\n")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max"]
for name in ("prof_gemm", "prof_attn", "prof_ln", "prof_conv"):
    p = f"{G}/{name}.ncu-rep"
    if not os.path.exists(p):
        continue
    txt = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    section(f"`ncu --set full --clock-control none` — {name}.ncu-rep")
    cols = [i for i, h in enumerate(hdr) if h in WANT]
    out.append("| kernel | grid | " + " | ".join(f"{hdr[i]} [{units[i]}]" for i in cols) + " |\n")
    out.append("|---|---|" + "---|" * len(cols) + "\n")
    for r in rows[2:]:
        out.append(f"| `{r[hdr.index('Kernel Name')][:60]}` | {r[hdr.index('Grid Size')]} | " + " | ".join(r[i] for i in cols) + " |\n")

if os.path.exists(f"{G}/bench_kernels.jsonl"):
    section("kernel micro-benchmarks (`tools/bench_kernels.py`, CUDA events, L2 flushed; library kernels for reference only)")
    out.append("| kernel | case | config | ms | TFLOP/s or GB/s |\n|---|---|---|---|---|\n")
    for line in open(f"{G}/bench_kernels.jsonl"):
        r = json.loads(line)
        if "error" in r:
            continue
        case = r.get("name") or (f"L={r['L']}" if "L" in r else f"rows={r.get('rows')}")
        cfg = f"bn={r['block_n']} cg={r['cta_group']}" if "block_n" in r else ""
        val = r.get("tflops", r.get("gbs", 0))
        out.append(f"| {r['kernel']} | {case} | {cfg} | {r['ms']:.4f} | {val:.1f} |\n")

# per-launch DRAM bytes of the captured launches -> profiles/<tag>_ncu_traffic.json (bench.py's roofline.traffic)
traffic = {}
for name, key in (("prof_gemm", "gemm"), ("prof_attn", "attention")):
    p = f"{G}/{name}.ncu-rep"
    if not os.path.exists(p):
        continue
    rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def val(r, n):
        i = hdr.index(n)
        return float(r[i].replace(",", "")) * mult.get(units[i], 1)
    L = [{"kernel": r[hdr.index("Kernel Name")][:70], "grid": r[hdr.index("Grid Size")],
          "dram_bytes": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"), "time_us": val(r, "gpu__time_duration.sum"),
          "tensor_pipe_active_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")} for r in rows[2:]]
    traffic[key] = {"launches": L, "avg_dram_bytes_per_launch": sum(x["dram_bytes"] for x in L) / len(L)}
if traffic:
    traffic["note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full --clock-control none` "
                       f"(profiles/{tag}_summary.md); cold caches")
    json.dump(traffic, open(f"profiles/{tag}_ncu_traffic.json", "w"), indent=1)

sp = [(w, f"profiles/{tag}_bench_sp{w}.json") for w in (2, 4, 8) if os.path.exists(f"profiles/{tag}_bench_sp{w}.json")]
if sp:
    section("one image over W GPUs: `bench.py --gpus W --mode sp` (sequence-parallel, strong scaling; rank 0's kernel times)")
    out.append("| W | images/s | ms per image | e2e images/s | gemm ms | attention ms | LayerNorm ms | barriers+other ms | SM MHz |\n|---|---|---|---|---|---|---|---|---|\n")
    for w, path in sp:
        b = json.load(open(path))
        k = b["kernel_time_share"]
        out.append(f"| {w} | {b['value']:.3f} | {b['ms_per_step']:.0f} | {b['e2e']['value']:.3f} | {k['gemm_ms']:.0f} | {k['attention_ms']:.0f} | "
                   f"{k['ln_modulate_ms']:.0f} | {k['other_ms']:.0f} | {b['clocks']['sm_mhz']:.0f} |\n")
    for w in (2, 4, 8):
        rp = f"profiles/{tag}_sp{w}_report.json"
        if os.path.exists(rp):
            out.append(f"\nparity report W={w} (`tests/sp_worker.py`): `{json.dumps(json.load(open(rp)))}`\n")

os.makedirs("profiles", exist_ok=True)
open(f"profiles/{tag}_summary.md", "w").write("".join(out))
print(f"wrote profiles/{tag}_summary.md ({sum(len(s) for s in out)} bytes)")
