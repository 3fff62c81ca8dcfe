"""Turns an ncu report (.ncu-rep, captured with --set full --import-source on) into the short text summary kept under
profiles/.   python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_elapsed.avg.per_second",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__sass_thread_inst_executed_op_fadd_pred_on.sum",
        "sm__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__warps_eligible.avg.per_cycle_active"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for k, row in enumerate(rows[2:]):
        name = row[hdr.index("Kernel Name")]
        print(f"== launch {k}: {name[:150]}")
        for key in KEYS:
            if key in hdr:
                i = hdr.index(key)
                print(f"   {key:75s} {row[i]:>16s} {units[i]}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    try:
        h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    except StopIteration:
        return
    hdr = rows[h]
    stalls = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    tot = {c: 0.0 for c in stalls}
    for r in rows[h + 1:]:
        if len(r) != len(hdr):
            continue
        for c in stalls:
            try:
                tot[c] += float(r[hdr.index(c)])
            except ValueError:
                pass
    s = sum(tot.values()) or 1.0
    print("== warp stall samples (source page, first launch)")
    for c, v in sorted(tot.items(), key=lambda x: -x[1])[:8]:
        print(f"   {c:28s} {100*v/s:5.1f} %")


if __name__ == "__main__":
    main(sys.argv[1])
