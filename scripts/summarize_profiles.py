"""Turns the ncu outputs in gpurun_out/ into the tracked summaries under profiles/ (run in the build container):
  launches CSV (ncu --metrics gpu__time_duration.sum)      -> per-kernel share of ONE denoising step
  GEMM / attention .ncu-rep (ncu --set full)               -> raw CSV + a table of the metrics DESIGN.md quotes
usage: python scripts/summarize_profiles.py <tag> <launches.csv> <gemm.ncu-rep> <attn.ncu-rep>"""
import csv, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]


def short(name):
    m = re.search(r"(mmada::[\w]+(<[^>]*>)?)", name)
    return m.group(1) if m else re.sub(r"\(.*", "", name)[-70:]


def launches(tag, path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    seq = [(short(r[4]), float(r[14].replace(",", "")) / 1e3) for r in rows]           # us
    marks = [i for i, (n, _) in enumerate(seq) if "t2i_sample" in n]
    a, b = marks[-2] + 1, marks[-1] + 1                                                 # one full denoising step
    step = seq[a:b]
    tot = sum(t for _, t in step)
    by = {}
    for n, t in step:
        c, s = by.get(n, (0, 0.0))
        by[n] = (c + 1, s + t)
    out = [f"one denoising step (launches between two t2i_sample kernels) = {tot / 1e3:.2f} ms of kernel time under ncu "
           f"(serialised, cold-cache); {len(step)} launches"]
    for n, (c, s) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{100 * s / tot:6.2f}%  {c:4d} launches  {s / c:9.1f} us avg  {n}")
    open(os.path.join(P, f"{tag}_launch_summary.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


def full(tag, what, rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(os.path.join(P, f"{tag}_{what}_ncu_full_raw.csv"), "w").write(raw)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    res = []
    for r in rows[2:]:
        d = {k: r[hdr.index(k)] for k in KEYS if k in hdr}
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):        # to bytes, whatever unit ncu picked for the column
            if k in hdr:
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[units[hdr.index(k)]]
                d[k + ".bytes"] = float(r[hdr.index(k)].replace(",", "")) * scale
        d["kernel"] = short(r[hdr.index("Kernel Name")])
        res.append(d)
        out.append(json.dumps(d))
    open(os.path.join(P, f"{tag}_{what}_ncu_full_summary.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))
    return res


def traffic(tag, res):
    """bench.py's roofline.traffic: DRAM bytes per GEMM launch, averaged over the launches of one denoising step
    (32 layers x {qkv, attn_out, gate_up, ff_out} + the lm_head), from the five captured launches."""
    M, d, f = 16 * 1539, 4096, 12288
    alg = {"qkv": 2 * (M * d + 3 * d * d + M * 3 * d), "attn_out": 2 * (M * d + d * d) + 4 * 2 * M * d + 2 * M * d,
           "gate_up": 2 * (M * d + 2 * f * d + M * f), "ff_out": 2 * (M * f + d * f) + 4 * 2 * M * d + 2 * M * d,
           "lm_head": 2 * (16384 * d + 8192 * d) + 4 * 16384 * 8192}
    shapes = {}
    for r in res:
        epi = int(re.search(r"gemm_kernel<\d+, (\d+)", r["kernel"]).group(1))
        ms = float(r["gpu__time_duration.sum"])
        ms = ms / 1e3 if ms > 50 else ms                                    # us or ms column
        name = {3: "gate_up", 7: "qkv", 1: "lm_head"}.get(epi) or ("ff_out" if ms > 1.0 else "attn_out")
        shapes[name] = {"epilogue": epi, "dram_bytes": r["dram__bytes_read.sum.bytes"] + r["dram__bytes_write.sum.bytes"],
                        "algorithmic_bytes": alg[name], "duration_ms": ms,
                        "tensor_pipe_active_pct": float(r["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"])}
    n = 32 * 4 + 1
    avg = (32 * sum(shapes[k]["dram_bytes"] for k in ("qkv", "attn_out", "gate_up", "ff_out")) + shapes["lm_head"]["dram_bytes"]) / n
    avg_alg = (32 * sum(alg[k] for k in ("qkv", "attn_out", "gate_up", "ff_out")) + alg["lm_head"]) / n
    out = {"dram_bytes_per_launch_avg": avg, "algorithmic_bytes_per_launch_avg": avg_alg, "per_shape": shapes,
           "source": f"ncu --set full (dram__bytes_read.sum + dram__bytes_write.sum) on bench.py --layers 2 --steps 2 --warmup 3, "
                     f"launches 24..28; profiles/{tag}_gemm_ncu_full_raw.csv; average over the {n} GEMM launches of one "
                     "denoising step (32 x the four block shapes + the lm_head)"}
    json.dump(out, open(os.path.join(P, f"{tag}_gemm_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    tag, lcsv, grep_, arep = sys.argv[1:5]
    launches(tag, lcsv)
    traffic(tag, full(tag, "gemm", grep_))
    full(tag, "attention_duo", arep)
