"""Condense ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.
usage: python profiles/summarize.py <round-prefix> <tag> [<tag> ...]
reads gpurun_out/launches_<tag>.csv and gpurun_out/prof_<tag>.ncu-rep, writes profiles/<round-prefix>_<tag>.md"""
import collections
import csv
import io
import os
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "smsp__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def main():
    prefix = sys.argv[1]
    for tag in sys.argv[2:]:
        out = [f"# ncu summary {tag}", ""]
        lc = f"gpurun_out/launches_{tag}.csv"
        if os.path.exists(lc):
            rows = [r for r in csv.reader(open(lc)) if len(r) > 5]
            hdr = rows[0]
            ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
            agg = collections.defaultdict(list)
            for r in rows[1:]:
                try:
                    agg[r[ki]].append(float(r[vi].replace(",", "")))
                except ValueError:
                    pass
            tot = sum(sum(v) for v in agg.values())
            out.append("## launch list (`--metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)")
            out.append("")
            out.append("| kernel | launches | mean us | share of GPU time |")
            out.append("|---|---|---|---|")
            for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
                out.append(f"| `{k[:110]}` | {len(v)} | {sum(v) / len(v) / 1000:.2f} | {sum(v) / tot:.3f} |")
            out.append("")
        rep = f"gpurun_out/prof_{tag}.ncu-rep"
        if os.path.exists(rep):
            raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
            rows = list(csv.reader(io.StringIO(raw)))
            hdr = rows[0]
            out.append("## `ncu --set full --clock-control none` of the dominant kernel (per launch; units as reported)")
            out.append("")
            kn = hdr.index("Kernel Name")
            out.append(f"kernel: `{rows[2][kn][:160]}`")
            out.append("")
            out.append("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(rows) - 2)) + " |")
            out.append("|---|---|" + "---|" * (len(rows) - 2))
            for i, h in enumerate(hdr):
                if h in KEYS or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")):
                    vals = [r[i] for r in rows[2:]]
                    try:
                        if max(float(v.replace(",", "")) for v in vals) == 0.0 and "stalled" in h:
                            continue
                    except ValueError:
                        pass
                    out.append(f"| {h} | {rows[1][i]} | " + " | ".join(vals) + " |")
            out.append("")
        path = f"profiles/{prefix}_{tag}.md"
        open(path, "w").write("\n".join(out) + "\n")
        print("wrote", path)


if __name__ == "__main__":
    main()
