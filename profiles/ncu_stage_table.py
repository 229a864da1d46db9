"""Turns `ncu -i <report> --page raw --csv` output into (a) a trimmed CSV with the columns the roofline numbers come from
and (b) a markdown table: per stage kernel duration, DRAM bytes, achieved HBM GB/s against the measured peak, and the
SM-side limiter counters.

    ncu -i gpurun_out/X.ncu-rep --page raw --csv > /tmp/raw.csv
    python profiles/ncu_stage_table.py /tmp/raw.csv profiles/ncu_rN_stages [frames_in_batch]
"""
import csv
import json
import os
import sys

KEEP = [
    "ID", "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def to_us(v, unit):
    return float(v) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}[unit]


def main():
    raw, out = sys.argv[1], sys.argv[2]
    frames = int(sys.argv[3]) if len(sys.argv) > 3 else None
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in KEEP if c in ix]
    with open(out + "_raw.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols)
        w.writerow([units[ix[c]] for c in cols])
        for r in data:
            w.writerow([r[ix[c]] for c in cols])
    peak = 6525.2
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    lines = [f"| kernel | grid x block | regs | time us | DRAM read MB | DRAM write MB | DRAM GB/s | % of {peak:.0f} GB/s | issue slots busy % | warps active % | top stall (per issue) |",
             "|---|---|---|---|---|---|---|---|---|---|---|"]
    tot_t = tot_b = 0.0
    for r in data:
        g = lambda k: r[ix[k]]                                    # noqa: E731
        u = lambda k: units[ix[k]]                                # noqa: E731
        t = to_us(g("gpu__time_duration.sum"), u("gpu__time_duration.sum"))
        rd = to_bytes(g("dram__bytes_read.sum"), u("dram__bytes_read.sum"))
        wr = to_bytes(g("dram__bytes_write.sum"), u("dram__bytes_write.sum"))
        gbs = (rd + wr) / (t * 1e-6) / 1e9
        stalls = {k.split("issue_stalled_")[1].split("_per_issue")[0]: float(g(k)) for k in cols if "issue_stalled" in k}
        top = max(stalls, key=stalls.get)
        name = g("Kernel Name").split("(")[0].replace("void ", "").replace("mmw::", "")
        lines.append(f"| `{name}` | {g('launch__grid_size')} x {g('launch__block_size')} | {g('launch__registers_per_thread')} | {t:.1f} | {rd / 1e6:.1f} | "
                     f"{wr / 1e6:.1f} | {gbs:.0f} | {100 * gbs / peak:.1f} | {float(g('smsp__issue_active.avg.pct_of_peak_sustained_active')):.1f} | "
                     f"{float(g('sm__warps_active.avg.pct_of_peak_sustained_active')):.1f} | {top} {stalls[top]:.2f} |")
        tot_t += t
        tot_b += rd + wr
    lines.append(f"| **sum** | | | {tot_t:.1f} | | | {tot_b / (tot_t * 1e-6) / 1e9:.0f} | {100 * tot_b / (tot_t * 1e-6) / 1e9 / peak:.1f} | | | |")
    if frames:
        lines.append("")
        lines.append(f"{frames} frames per launch: {tot_b / frames / 1e6:.2f} MB of DRAM traffic per frame, {tot_t / frames:.2f} us per frame under ncu "
                     "(per-launch times under ncu are serialised and cold-cache; bench.py's CUDA-event times are the ones to quote).")
    open(out + ".md", "a").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
