"""Turn an ncu report (gpurun_out/*.ncu-rep, scratch) into the small per-kernel CSV kept under profiles/:
    python profiles/summarize.py gpurun_out/prof_tc_r1d.ncu-rep profiles/r1_ncu_tc_gemm_full.csv
Needs the `ncu` CLI (no GPU): ncu -i <rep> --page raw --csv."""
import csv, io, subprocess, sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
    "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]

def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    names = [r[hdr.index("Kernel Name")] for r in body]
    short = []
    for nm in names:
        nm = nm.split("::")[-1].split("(")[0]
        short.append(nm)
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + short)
        for m in METRICS:
            if m not in hdr:
                continue
            i = hdr.index(m)
            w.writerow([m, units[i]] + [r[i] for r in body])
    print(open(out).read())

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
