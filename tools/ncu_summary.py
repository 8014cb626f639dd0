"""Condense `ncu -i X.ncu-rep --page raw --csv` into the handful of numbers the roofline discussion needs.
Usage: python tools/ncu_summary.py raw.csv [raw2.csv ...]  -> text on stdout (commit under profiles/)."""
import csv
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (of active cycles)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (of elapsed)"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__occupancy_limit_registers", "occ. limit regs (blocks)"),
    ("launch__occupancy_limit_shared_mem", "occ. limit smem (blocks)"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_throttle / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
]

for fn in sys.argv[1:]:
    rows = list(csv.reader(open(fn)))
    if len(rows) < 3:
        print("%s: empty" % fn)
        continue
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# %s" % fn)
    for r in rows[2:]:
        print("kernel %s   grid %s block %s" % (r[idx["Kernel Name"]][:80], r[idx.get("Grid Size", 0)], r[idx.get("Block Size", 0)]))
        for key, label in WANT:
            if key in idx:
                print("    %-44s %s %s" % (label, r[idx[key]], units[idx[key]]))
        try:
            rd = float(r[idx["dram__bytes_read.sum"]].replace(",", "")); wr = float(r[idx["dram__bytes_write.sum"]].replace(",", ""))
            mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd *= mul.get(units[idx["dram__bytes_read.sum"]], 1); wr *= mul.get(units[idx["dram__bytes_write.sum"]], 1)
            dur = float(r[idx["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(units[idx["gpu__time_duration.sum"]], 1e-9)
            print("    %-44s %.1f MB -> %.0f GB/s under the profiler" % ("dram traffic (read+write)", (rd + wr) / 1e6, (rd + wr) / dur / 1e9))
        except Exception:
            pass
