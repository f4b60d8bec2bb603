"""Turns an .ncu-rep (ncu --set full) into the short text summary committed under profiles/.
Usage: python profiles/summarize.py gpurun_out/X.ncu-rep [--json profiles/X.json --kernel mppi_tick_kernel --K 1048576 --T 50] > profiles/X.txt
With --json the counters bench.py quotes (DRAM bytes and executed warp-instructions per launch of the named kernel) are
written next to the text summary together with the commit they were measured at, so the bench line never carries a literal
copied from an old capture."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]


def _opt(name, default=None):
    return sys.argv[sys.argv.index(name) + 1] if name in sys.argv else default


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print("# ncu --set full --clock-control none summary of", rep.split("/")[-1])
    for r in data:
        print("\n## kernel:", r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                print("%-85s %-14s %s" % (k, units[hdr.index(k)], r[hdr.index(k)]))
        print("# warp stall reasons (average warps stalled per issue-active cycle)")
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
                v = float(r[i] or 0)
                if v >= 0.05:
                    print("  %-40s %.3f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))


    js = _opt("--json")
    if js:
        import json
        kname = _opt("--kernel", "mppi_tick_kernel")
        rows_k = [r for r in data if kname in r[hdr.index("Kernel Name")]]
        if rows_k:
            def col(r, k):
                return float(r[hdr.index(k)].replace(",", "")) if k in hdr and r[hdr.index(k)] else None

            def to_bytes(r, k):
                v, u = col(r, k), units[hdr.index(k)].lower() if k in hdr else ""
                return None if v is None else v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            r = rows_k[-1]
            commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
            out_js = {"kernel": r[hdr.index("Kernel Name")], "source": "ncu --set full --clock-control none, " + rep.split("/")[-1] + ", commit " + commit,
                      "commit": commit, "K": int(_opt("--K", "1048576")), "T": int(_opt("--T", "50")),
                      "dram_bytes_per_launch": (to_bytes(r, "dram__bytes_read.sum") or 0) + (to_bytes(r, "dram__bytes_write.sum") or 0),
                      "inst_executed_per_launch": col(r, "smsp__inst_executed.sum"),
                      "issue_active_pct": col(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                      "duration_us_under_ncu": col(r, "gpu__time_duration.sum")}
            json.dump(out_js, open(js, "w"), indent=1)


if __name__ == "__main__":
    main()
