"""Markdown summary of an `ncu --set full` report (read here, no GPU needed):
python scripts/summarize_ncu.py gpurun_out/r1c_ncu.ncu-rep > profiles/r1c_ncu_summary.md
(or of its raw page exported as CSV on the GPU box: ... gpurun_out/r2i_ncu_factor_raw.csv)"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    if rep.endswith(".csv"):     # raw page already exported on the GPU box (ncu -i X.ncu-rep --page raw --csv)
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full summary of `{rep}`\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"## {d.get('Kernel Name', '?')[:90]}  (launch id {d.get('ID', '?')}, grid {d.get('launch__grid_size', '?')})\n")
        print("| metric | value | unit |\n|---|---|---|")
        for m in METRICS:
            if m in d:
                print(f"| {m} | {d[m]} | {u[m]} |")
        st = sorted(((float(v.replace(',', '')), k[len(STALL):].replace('_per_issue_active.ratio', ''))
                     for k, v in d.items() if k.startswith(STALL) and k.endswith("per_issue_active.ratio") and v),
                    reverse=True)[:6]
        if st:
            print("\nTop stall reasons (warps per issue-active cycle):\n")
            for v, k in st:
                print(f"- {k}: {v:.2f}")
        print()


if __name__ == "__main__":
    main()
