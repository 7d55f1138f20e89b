"""Summarise an .ncu-rep (ncu --set full) into a small CSV: one row per captured launch, the metrics that the
roofline discussion uses.  Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xyz.csv"""
import csv
import io
import subprocess
import sys

WANT = [
    ("Kernel Name", "kernel"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe_alu_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe_fma_pct"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "pipe_fp64_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe_lsu_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "pipe_tensor_pct"),
    ("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "inst_tensor_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("smsp__inst_executed.sum", "warp_insts"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    tensor_like = [h for h in hdr if "tensor" in h and "pct" in h]
    out = csv.writer(sys.stdout)
    cols = [(h, n) for h, n in WANT if h in idx]
    extra = [h for h in tensor_like if h not in dict(cols)][:4]
    out.writerow([n + (" [%s]" % units[idx[h]] if units[idx[h]] else "") for h, n in cols] + extra)
    for r in rows[2:]:
        out.writerow([r[idx[h]] for h, _ in cols] + [r[idx[h]] for h in extra])


if __name__ == "__main__":
    main(sys.argv[1])
