"""Per-CUDA-source-line instruction and stall-sample shares of one kernel in an .ncu-rep
(ncu -i REP --page source --print-source cuda,sass --csv --kernel-name K  piped in as a file)."""
import csv
import sys


def main(path, thresh=0.008):
  rows = list(csv.reader(open(path, errors="replace")))
  hi = next(i for i, r in enumerate(rows) if len(r) > 8 and r[0] == "Line No")
  hdr = rows[hi]
  ie, sm = hdr.index("Instructions Executed"), hdr.index("# Samples")
  out, tot, tsm = [], 0, 0
  for r in rows[hi + 1:]:
    if len(r) <= ie or r[2] != "-":   # per-line summary rows have '-' as the address
      continue
    try:
      v, s = int(r[ie]), int(r[sm])
    except ValueError:
      continue
    tot += v
    tsm += s
    out.append((v, s, r[0], r[1][:120]))
  print("total warp instructions %d, samples %d" % (tot, tsm))
  for v, s, ln, src in out:
    if v > tot * thresh or s > tsm * thresh:
      print("%5.1f%% inst %5.1f%% smp  L%-5s | %s" % (100.0 * v / tot, 100.0 * s / max(tsm, 1), ln, src))


if __name__ == "__main__":
  main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.008)
