"""Text summary of an `ncu --set full` report for profiles/: per kernel launch the duration, DRAM traffic,
pipe utilisation and the top warp-stall reasons / source lines.
  python tools/ncu_summary.py rep.ncu-rep [kernel-regex] > profiles/rNN_<name>.summary.txt"""
import csv
import subprocess
import sys

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else None
KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
print(f"# {rep}")
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    print("-" * 100)
    for k in KEYS:
        if k in d and d[k] != "":
            print(f"{k:85s} {d[k]} {u.get(k, '')}")
    tm = [h for h in hdr if "tensor" in h and "pct_of_peak_sustained_elapsed" in h and d.get(h) not in ("", "0", None)]
    for h in tm[:6]:
        if h not in KEYS:
            print(f"{h:85s} {d[h]} {u.get(h, '')}")
sys.stdout.flush()
cmd = [sys.executable, __file__.replace("ncu_summary.py", "src_stalls.py"), rep] + ([kre] if kre else [])
print("-" * 100)
print(subprocess.run(cmd + ([] if kre else []) + (["18"] if kre else []), capture_output=True, text=True).stdout)
