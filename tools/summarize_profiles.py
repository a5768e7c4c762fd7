#!/usr/bin/env python
"""Turn gpurun_out/ artefacts into the tracked summaries under profiles/:
  launches csv (ncu --metrics gpu__time_duration.sum) -> per-kernel share table
  *.ncu-rep (ncu --set full)                           -> key raw metrics per captured launch
Usage: python tools/summarize_profiles.py <tag> [launches.csv] [report.ncu-rep ...]"""
import collections, csv, json, os, re, subprocess, sys


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        k = re.sub(r"\(.*", "", row["Kernel Name"])
        agg[k][0] += 1; agg[k][1] += v; tot += v
    out = ["| kernel | launches | total ms | avg us | share |", "|---|---:|---:|---:|---:|"]
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:30]:
        out.append(f"| `{k}` | {n} | {t / 1e6:.3f} | {t / n / 1e3:.2f} | {100 * t / tot:.1f}% |")
    return "\n".join(out), tot / 1e6


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]


def report(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    rows = [r for r in rows if len(r) > 10]
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": re.sub(r"\(.*", "", r[hdr.index("Kernel Name")])}
        for w in WANT:
            if w in hdr:
                d[w] = f"{r[hdr.index(w)]} {units[hdr.index(w)]}".strip()
        out.append(d)
    return out


if __name__ == "__main__":
    tag = sys.argv[1]
    os.makedirs("profiles", exist_ok=True)
    for p in sys.argv[2:]:
        if p.endswith(".csv"):
            table, tot = launches(p)
            open(f"profiles/{tag}_launches.md", "w").write(f"# {tag}: ncu launch list summary ({os.path.basename(p)}, total {tot:.1f} ms)\n\n" + table + "\n")
        else:
            json.dump(report(p), open(f"profiles/{tag}_{os.path.basename(p).replace('.ncu-rep', '')}.json", "w"), indent=1)
    print("ok")
