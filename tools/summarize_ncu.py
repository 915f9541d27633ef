"""Turn ncu reports (gpurun_out/*.ncu-rep, launches.csv) into the small text summaries kept under profiles/."""
import csv, io, subprocess, sys
from collections import defaultdict
from pathlib import Path

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]


EXTRA = ["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
         "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
         "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__mem_tensor_writes_op_stt.sum.pct_of_peak_sustained_elapsed",
         "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
         "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
         "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
         "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
         "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def summarize(rep, title):
    hdr, units, rows = raw(rep)
    lines = [f"# {title}", f"# source: {Path(rep).name} (ncu --set full --clock-control none --import-source on; one launch)"]
    for r in rows:
        lines.append(f"kernel: {r[hdr.index('Kernel Name')]}")
        for k in hdr:
            if k in KEYS or k in EXTRA:
                i = hdr.index(k)
                lines.append(f"  {k:80s} {r[i]:>18s} {units[i]}")
    return "\n".join(lines) + "\n"


def launches(csvpath):
    rows = [r for r in csv.reader(open(csvpath)) if len(r) > 5]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        agg[r[ik]][0] += 1
        agg[r[ik]][1] += v
    tot = sum(v[1] for v in agg.values())
    lines = ["# per-kernel share of the captured region (gpu__time_duration.sum, cold-cache & serialised: compare shares, not absolutes)"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{100 * t / tot:6.2f} %  {n:5d} launches  {t / 1e3:10.1f} us total  {t / n / 1e3:8.2f} us avg  {k[:110]}")
    return "\n".join(lines) + "\n"


if __name__ == "__main__":
    out = Path("profiles")
    out.mkdir(exist_ok=True)
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    g = Path("gpurun_out")
    if (g / "prof_decode_gate_up.ncu-rep").exists():
        (out / f"{tag}_decode_gate_up_m1_ncu.txt").write_text(summarize(g / "prof_decode_gate_up.ncu-rep", "fused small-M kernel, gate_up 4096->28672 (P=2), M=1"))
    if (g / "prof_gemm_gate_up.ncu-rep").exists():
        (out / f"{tag}_gemm_gate_up_m4096_ncu.txt").write_text(summarize(g / "prof_gemm_gate_up.ncu-rep", "tcgen05 INT4-dequant GEMM, gate_up 4096->28672, M=4096"))
    if (g / "launches.csv").exists():
        (out / f"{tag}_bench_launches.txt").write_text(launches(g / "launches.csv"))
