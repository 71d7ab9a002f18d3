"""Summarise one `ncu --set full --import-source on` capture for profiles/.

    python tools/ncu_summary.py gpurun_out/melif_r02b.ncu-rep profiles/r01_melif_v6_r02b_ncu_summary.csv

Writes the rows of `--page raw` that profiles/README.md quotes, followed by the warp-stall
samples of `--page source` (SASS view) aggregated per stall reason and per block of
instructions, and the instructions with the most samples.  The per-block table is what
located the divergent item 0 of the polar step and the spilled band weights.
"""
import csv
import subprocess
import sys

RAW_KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "sm__cycles_elapsed.max",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]
STALLS = ["stall_long_sb", "stall_short_sb", "stall_barrier", "stall_wait", "stall_mio",
          "stall_not_selected", "stall_selected", "stall_no_inst", "stall_math",
          "stall_branch_resolving", "stall_dispatch", "stall_lg"]


def ncu_csv(report, *args):
    out = subprocess.run(["ncu", "-i", report, "--csv", *args], capture_output=True, text=True,
                         check=True).stdout
    return list(csv.reader(out.splitlines()))


def main(report, dest, block=300, top=25):
    lines = []
    raw = ncu_csv(report, "--page", "raw")
    head, units, vals = raw[0], raw[1], raw[-1]
    lines.append(f'kernel,"{vals[head.index("Kernel Name")][:90]}"')
    for key in RAW_KEYS:
        if key in head:
            i = head.index(key)
            lines.append(f"{key},{units[i]},{vals[i]}")

    sass = ncu_csv(report, "--page", "source", "--print-source", "sass")
    hdr, data = sass[1], sass[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def num(row, key):
        try:
            return float(row[col[key]])
        except (ValueError, IndexError, KeyError):
            return 0.0

    total = sum(num(r, "# Samples") for r in data) or 1.0
    total_inst = sum(num(r, "Instructions Executed") for r in data) or 1.0
    lines.append(f"warp_stall_samples,count,{int(total)}")
    for s in STALLS:
        lines.append(f"{s},fraction_of_samples,{sum(num(r, s) for r in data) / total:.4f}")
    lines.append("# per block of SASS instructions: first index, % samples, % warp instructions, "
                 "% of all samples stalled on long_sb / short_sb / barrier / wait / mio, "
                 "shared wavefronts actual / ideal (millions)")
    for start in range(0, len(data), block):
        seg = data[start:start + block]
        pct = lambda k: 100.0 * sum(num(r, k) for r in seg) / total
        lines.append("block,%d,%.1f,%.1f,%.1f,%.1f,%.1f,%.1f,%.1f,%.2f,%.2f" % (
            start, pct("# Samples"), 100.0 * sum(num(r, "Instructions Executed") for r in seg) / total_inst,
            pct("stall_long_sb"), pct("stall_short_sb"), pct("stall_barrier"), pct("stall_wait"),
            pct("stall_mio"), sum(num(r, "L1 Wavefronts Shared") for r in seg) / 1e6,
            sum(num(r, "L1 Wavefronts Shared Ideal") for r in seg) / 1e6))
    lines.append("# instructions with the most samples: index, SASS, samples, dominant stall")
    order = sorted(range(len(data)), key=lambda i: -num(data[i], "# Samples"))[:top]
    for i in sorted(order):
        reasons = {s: num(data[i], s) for s in STALLS}
        worst = max(reasons, key=reasons.get)
        lines.append('hot,%d,"%s",%d,%s' % (i, data[i][1].strip()[:70], num(data[i], "# Samples"), worst))
    with open(dest, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
