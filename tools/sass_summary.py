"""Per-kernel SASS evidence for profiles/: for every kernel of libodin_b200.so the instruction count and the counts of
the mnemonics that prove which hardware path it uses (tcgen05: UTCHMMA / UTCQMMA, TMEM loads / stores: LDTM / STTM,
TMA / bulk copies: UBLKCP / UTMALDG, mbarrier: SYNCS, packed fp32: FFMA2 / FADD2 / FMUL2, async copies: LDGSTS, ...).

    python tools/sass_summary.py odin_b200/lib/libodin_b200.so > profiles/r02_sass_summary.txt"""
import collections
import re
import subprocess
import sys

KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "DMMA", "FFMA2", "FADD2", "FMUL2",
        "FFMA", "DFMA", "HMMA", "LDS", "STS", "SHFL", "RED", "ATOM", "BAR", "MUFU"]


def main(path):
  txt = subprocess.run(["cuobjdump", "-sass", path], stdout=subprocess.PIPE, text=True).stdout
  txt = subprocess.run(["cu++filt"], input=txt, stdout=subprocess.PIPE, text=True).stdout
  cur, kernels = None, collections.OrderedDict()
  for line in txt.splitlines():
    m = re.search(r"Function : (.*)", line)
    if m:
      cur = m.group(1).strip()
      kernels[cur] = collections.Counter()
      continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
      op = m.group(2)
      kernels[cur]["_total"] += 1
      for k in KEYS:
        if op == k or op.startswith(k + "."):
          kernels[cur][k] += 1
  print("# SASS summary of %s (cuobjdump -sass; static counts per kernel; only non-zero columns shown)" % path)
  for name, c in kernels.items():
    short = re.sub(r"\(.*\)$", "", name.replace("odin::", ""))
    cols = " ".join("%s=%d" % (k, c[k]) for k in KEYS if c[k])
    print("%-92s insts=%-6d %s" % (short[:92], c["_total"], cols))


if __name__ == "__main__":
  main(sys.argv[1] if len(sys.argv) > 1 else "odin_b200/lib/libodin_b200.so")
