"""Condense `ncu -i X.ncu-rep --page raw --csv` into the few metrics the design notes cite.
usage: python scripts/summarize_ncu.py raw.csv > profiles/summary.md"""
import csv
import sys

WANT = [
  ("gpu__time_duration.sum", "duration"),
  ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
  ("launch__waves_per_multiprocessor", "waves/SM"),
  ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
  ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
  ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
  ("smsp__inst_executed.sum", "warp instructions"),
  ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
  ("sm__inst_executed_pipe_fma.sum", "FMA-pipe instr"), ("sm__inst_executed_pipe_alu.sum", "ALU-pipe instr"),
  ("sm__inst_executed_pipe_xu.sum", "XU (MUFU) instr"), ("sm__inst_executed_pipe_lsu.sum", "LSU instr"),
  ("sm__inst_executed_pipe_fp64.sum", "FP64 instr"),
  ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
  ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
  ("lts__t_bytes.sum", "L2 bytes"),
  ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
  ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
  ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard"),
  ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard"),
  ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math pipe throttle"),
  ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall MIO throttle"),
  ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
  ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not selected"),
]

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
  print(f"### {r[idx['Kernel Name']]}\n")
  print("| metric | value |\n|---|---|")
  for key, label in WANT:
    if key in idx and r[idx[key]] != "":
      print(f"| {label} (`{key}`) | {r[idx[key]]} {units[idx[key]]} |")
  print()
