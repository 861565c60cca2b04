"""Key metrics of every kernel in an .ncu-rep (via `ncu -i REP --page raw --csv`), one column per kernel,
in the form committed under profiles/ (r01_*_keymetrics.csv)."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
]


def main(rep, dst):
  raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
  rows = list(csv.reader(raw.splitlines()))
  hdr, units, data = rows[0], rows[1], rows[2:]
  stalls = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
  with open(dst, "w") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + ["%s:%s" % (r[hdr.index("ID")], r[hdr.index("Kernel Name")][:60]) for r in data])
    for k in KEYS + stalls:
      if k in hdr:
        i = hdr.index(k)
        w.writerow([k, units[i]] + [r[i] for r in data])


if __name__ == "__main__":
  main(sys.argv[1], sys.argv[2])
