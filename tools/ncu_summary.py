"""Summarise an .ncu-rep (raw page) into the handful of metrics the roofline needs; prints a markdown table."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
def col(name):
    return hdr.index(name) if name in hdr else None
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "bf16 MMA ops % of peak (elapsed)"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"), ("lts__t_bytes.sum", "L2 bytes"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("smsp__inst_executed.sum", "inst"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %")]
idx = [(col(a), b) for a, b in want if col(a) is not None]
print("| " + " | ".join(b for _, b in idx) + " |")
print("|" + "---|" * len(idx))
for r in data:
    cells = []
    for i, b in idx:
        v = r[i]
        if b == "kernel":
            v = v.split("(")[0].replace("void ", "")[:60]
        else:
            v = v + " " + units[i]
        cells.append(v)
    print("| " + " | ".join(cells) + " |")
