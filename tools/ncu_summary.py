#!/usr/bin/env python
"""Key metrics of every kernel in an .ncu-rep (ncu -i <rep> --page raw --csv), one block each.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg",
    "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
]
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
print("# %s" % rep)
for r in rows[2:]:
  print("== %s" % r[col["Kernel Name"]][:110])
  for k in KEYS:
    hits = [h for h in hdr if h.endswith(k)]
    for h in hits[:1]:
      print("   %-82s %s %s" % (k, r[col[h]], units[col[h]]))
  # stall breakdown (cycles a warp spends stalled per issued instruction, by reason)
  stalls = [(h, r[col[h]]) for h in hdr if "average_warp_latency_issue_stalled" in h or
            "average_warps_issue_stalled" in h]
  def num(v):
    try:
      return float(v.replace(",", ""))
    except Exception:
      return 0.0
  for h, v in sorted(stalls, key=lambda hv: -num(hv[1]))[:8]:
    print("   stall %-76s %s" % (h.split("stalled_")[-1], v))
