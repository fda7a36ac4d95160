"""Summarise `ncu --set full` reports (gpurun_out/prof_*.ncu-rep) into one text table for profiles/.
    python scripts/ncu_summary.py gpurun_out/prof_attn_b2.ncu-rep ... > profiles/rN_ncu_full_kernels_summary.txt
"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]


def rows(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr = rd[0]
    units = rd[1]
    for r in rd[2:]:
        yield dict(zip(hdr, r)), dict(zip(hdr, units))


for path in sys.argv[1:]:
    print(f"== {path}")
    for r, u in rows(path):
        name = r.get("Kernel Name", "?")
        print(f"kernel {name}  grid {r.get('Grid Size')} block {r.get('Block Size')}")
        for k in WANT:
            hit = [h for h in r if h == k]
            if hit:
                print(f"    {k:75s} {r[k]:>18s} {u[k]}")
