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
    hdr = rows[0]
    out = []
    res = []
    for r in rows[2:]:
        d = {k: r[hdr.index(k)] for k in KEYS if k in hdr}
        d["kernel"] = short(r[hdr.index("Kernel Name")])
        res.append(d)
        out.append(json.dumps(d))
    open(os.path.join(P, f"{tag}_{what}_ncu_full_summary.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))
    return res


if __name__ == "__main__":
    tag, lcsv, grep_, arep = sys.argv[1:5]
    launches(tag, lcsv)
    full(tag, "gemm", grep_)
    full(tag, "attention_pair", arep)
