#!/usr/bin/env python3
"""profiles/sass_blocks.py <report.ncu-rep> <kernel regex> -- groups the SASS of one profiled kernel into runs of
instructions with the same execution count (~ basic blocks) and prints the heaviest ones with their warp-stall
samples: a quick "where do the instructions go" view when the CUDA-source page is not exported as CSV."""
import collections
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = next(r for r in rows if "Instructions Executed" in r)
ie, si, ss = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
tot, blocks, cur, addr0 = 0, [], None, None
for r in rows:
    if len(r) <= ie or r is hdr or not r[ie].isdigit():
        continue
    n = int(r[ie])
    tot += n
    if cur is None or cur[0] != n:
        cur = [n, 0, 0, [], r[0]]
        blocks.append(cur)
    cur[1] += 1
    cur[2] += int(r[ss])
    cur[3].append(r[si].strip().split()[0])
samples = sum(b[2] for b in blocks)
print("total warp instructions %d, stall samples %d" % (tot, samples))
for b in sorted(blocks, key=lambda b: -b[0] * b[1])[:top]:
    c = collections.Counter(b[3])
    print("%s exec %9d x %4d instr = %10d (%4.1f%%)  samples %5.1f%%  %s" % (
        b[4][-5:], b[0], b[1], b[0] * b[1], 100.0 * b[0] * b[1] / tot, 100.0 * b[2] / max(samples, 1), dict(c.most_common(7))))
