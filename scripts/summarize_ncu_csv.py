"""Summaries of `ncu --set full` captures that were exported to CSV on the GPU box (scripts/gpu_call*.sh:
`--page raw --csv` and `--page source --csv`): one markdown table of the key counters per kernel, the
stall-reason mix, and a JSON of the measured DRAM traffic per launch (profiles/ncu_traffic.json).

    python scripts/summarize_ncu_csv.py <prefix> <out.md> [traffic.json]      # prefix e.g. gpurun_out/c3
"""
import collections
import csv
import glob
import gzip
import json
import os
import sys

WANT = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs/thread"), ("launch__shared_mem_per_block_dynamic", "dyn smem/CTA"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
        ("smsp__inst_executed.sum", "warp instructions"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %")]


# prof_target.py target -> the bench workload whose shape it reproduces
WORKLOAD = {"stft_scm2": "cfg2", "stft_scm1": "cfg2", "stft": "cfg2", "filter_dual": "cfg2", "solve4": "cfg2", "istft": "cfg2",
            "stft_scm_c8": "cfg4_512", "stft_scm_c8_256": "cfg4_256", "masked_scm_zf8": "cfg4_512", "solve8": "cfg4_512",
            "tango_mid44": "cfg3", "filter_multi44": "cfg3", "tango_mid28": "cfg5", "filter_multi28": "cfg5"}


def raw(path):
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        return None
    hdr, units, vals = rows[0], rows[1], rows[-1]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def stalls(path):
    rows = list(csv.reader(gzip.open(path, "rt")))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    n = 0
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        n += int(r[ix["# Samples"]] or 0)
        for s in names:
            tot[s[6:]] += int(r[ix[s]] or 0)
    return n, tot


def main():
    prefix, out = sys.argv[1], sys.argv[2]
    traffic_path = sys.argv[3] if len(sys.argv) > 3 else None
    lines = ["# ncu --set full summaries (round 2)", "",
             "Captured with `scripts/prof_target.py <target>` (one kernel, BASELINE shapes, third launch), "
             "`--clock-control none`; exported to CSV on the GPU box.", ""]
    traffic = {}
    for path in sorted(glob.glob(prefix + "_raw_*.csv")):
        tag = os.path.basename(path)[len(os.path.basename(prefix)) + 5:-4]
        d = raw(path)
        if not d:
            continue
        name = d.get("Kernel Name", ("?", ""))[0]
        lines += ["## %s — `%s`" % (tag, name), "", "| counter | value |", "|---|---|"]
        for key, label in WANT:
            if key in d:
                v, u = d[key]
                lines.append("| %s | %s %s |" % (label, v, u))
        src = prefix + "_src_" + tag + ".csv.gz"
        if os.path.exists(src):
            n, tot = stalls(src)
            if n:
                lines.append("| stall mix (warp samples) | %s |" % ", ".join("%s %.0f%%" % (k, 100.0 * v / n) for k, v in tot.most_common(7)))
        lines.append("")
        try:
            def bts(key):
                v, u = d[key]
                mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                return float(v.replace(",", "")) * mult
            traffic[tag] = {"kernel": name, "workload": WORKLOAD.get(tag), "dram_bytes_per_launch": bts("dram__bytes_read.sum") + bts("dram__bytes_write.sum"),
                            "dram_read": bts("dram__bytes_read.sum"), "dram_write": bts("dram__bytes_write.sum"),
                            "duration_us_under_ncu": float(d["gpu__time_duration.sum"][0].replace(",", ""))}
        except Exception:
            pass
    open(out, "w").write("\n".join(lines) + "\n")
    if traffic_path:
        json.dump(traffic, open(traffic_path, "w"), indent=1)
    print(out, len(traffic), "kernels")


if __name__ == "__main__":
    main()
