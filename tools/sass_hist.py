"""Opcode histogram of one kernel of a cubin / object file (cuobjdump -sass), e.g.
   python tools/sass_hist.py odin_b200/csrc/.obj/fe_frame5.o 'fe_frame5_kernel<1024, short, 13, true>'
Static counts: every instruction once (loops are not weighted)."""
import collections
import re
import subprocess
import sys


def main():
  obj, pat = sys.argv[1], sys.argv[2]
  txt = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
  names = subprocess.run(["cu++filt"], input=txt, stdout=subprocess.PIPE, text=True).stdout
  cur, hist, total = None, collections.Counter(), 0
  for line in names.splitlines():
    m = re.search(r"Function : (.*)", line)
    if m:
      cur = m.group(1)
      continue
    if cur is None or pat not in cur:
      continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
      hist[m.group(2)] += 1
      total += 1
  print("kernel pattern %r: %d instructions" % (pat, total))
  for op, n in hist.most_common(40):
    print("%6d  %s" % (n, op))


if __name__ == "__main__":
  main()
