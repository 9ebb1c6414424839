"""Turn the ncu artefacts a gpurun call brought back into small text summaries under profiles/.

    python scripts/summarize_profile.py <launches.csv> <prof.ncu-rep> <tag>
"""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    cols, data = rows[hdr], rows[hdr + 1:]
    ki, vi, gi, bi = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Grid Size"), cols.index("Block Size")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0]
        agg.setdefault((name, r[gi], r[bi]), []).append(float(r[vi].replace(",", "")))
    ours = lambda name: not name.startswith("void at::")
    tot = sum(sum(v) for k, v in agg.items() if ours(k[0]))
    out = ["| kernel | grid | block | launches | mean us | share of disco kernels |", "|---|---|---|---|---|---|"]
    for (name, g, b), v in agg.items():
        share = sum(v) / tot if ours(name) else float("nan")
        out.append("| `%s` | %s | %s | %d | %.1f | %.3f |" % (name, g, b, len(v), sum(v) / len(v) / 1e3, share))
    return "\n".join(out)


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def raw(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        out.append("kernel: %s" % name)
        for i, h in enumerate(hdr):
            if h in WANT:
                out.append("  %-70s %s %s" % (h, vals[i], units[i]))
    return "\n".join(out)


if __name__ == "__main__":
    lcsv, rep, tag = sys.argv[1:4]
    print("# ncu summary %s\n\n## launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised)\n" % tag)
    print(launches(lcsv))
    print("\n## ncu --set full of the fused STFT+SCM kernel\n\n```\n%s\n```" % raw(rep))
