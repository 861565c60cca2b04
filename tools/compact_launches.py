"""Compacts an `ncu --metrics gpu__time_duration.sum --csv` log into id,kernel,grid,block,ns rows and
appends the per-kernel share table (the form committed under profiles/)."""
import collections
import csv
import re
import sys


def short(name):
  name = re.sub(r"\(.*", "", name)          # drop the argument list
  name = re.sub(r"<unnamed>::", "", name)
  return name[:90]


def main(src, dst):
  rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) >= 15 and r[0].isdigit()]
  agg = collections.OrderedDict()
  with open(dst, "w") as f:
    f.write("id,kernel,grid,block,gpu__time_duration.sum,unit\n")
    for r in rows:
      k = short(r[4])
      f.write('%s,"%s","%s","%s",%s,%s\n' % (r[0], k, r[8], r[7], r[14], r[13]))
      a = agg.setdefault(k, [0, 0.0])
      a[0] += 1
      a[1] += float(r[14].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    f.write("\n# share of the captured launches (cold-cache, serialised: compare shares, not absolutes)\n")
    f.write("# kernel,launches,total_ns,share\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
      f.write('# "%s",%d,%.0f,%.4f\n' % (k, v[0], v[1], v[1] / tot))


if __name__ == "__main__":
  main(sys.argv[1], sys.argv[2])
