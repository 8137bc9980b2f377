"""Condense an .ncu-rep (read here, no GPU needed) into the handful of metrics the roofline argument uses.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv, subprocess, sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__cycles_active.avg",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


_BYTE_UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def to_bytes(value, unit):
    """ncu scales every column on its own (a read in Gbyte next to a write in Mbyte): convert before adding."""
    return float(value.replace(",", "")) * _BYTE_UNITS[unit]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("# %s" % path)
    for r in rows[2:]:
        print("kernel: %s" % r[hdr.index("Kernel Name")][:110])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("  %-100s %16s %s" % (w, r[i], units[i]))
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        total = to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw])
        print("  traffic (dram read+write)  %.4f GB" % (total / 1e9))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
