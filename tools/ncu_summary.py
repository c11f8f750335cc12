"""Per-kernel summary of an `ncu --set full` report: python tools/ncu_summary.py report.ncu-rep  (reads it with ncu --page raw --csv)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(hdr)}
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (sm__pipe_tensor_cycles_active)"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
        ("sm__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts % of peak"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "registers / thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__inst_executed.sum", "warp instructions"),
        ("lts__t_bytes.sum", "L2 bytes"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "memory throughput %")]
seen = {}
for r in data:
    name = r[col["Kernel Name"]]
    k = seen.get(name, 0)
    seen[name] = k + 1
    print(f"== {name[:110]}   (capture {k})")
    for m, label in want:
        if m in col:
            print(f"   {label:58s} {r[col[m]]:>18s} {units[col[m]]}")
    print()
