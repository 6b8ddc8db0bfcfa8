#!/usr/bin/env python
"""Dynamic instruction census from `ncu --page source --csv`: per opcode executed warp-instructions and average lanes,
and the instruction ranges (by executed count plateau) -- where the issue slots of a kernel go.
usage: src_hot.py source.csv [ray_steps]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
steps = float(sys.argv[2]) if len(sys.argv) > 2 else None
tot = 0; thr = 0
by = collections.Counter(); byl = collections.Counter()
data = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    n = int(r[ix["Instructions Executed"]]); t = int(r[ix["Thread Instructions Executed"]])
    tot += n; thr += t; by[op] += n; byl[op] += t
    data.append((n, t, src))
print(f"total warp-instr {tot:.4e}  lanes/instr {thr / tot:.2f}" + (f"  per warp-step(32 ray.steps) {tot / (steps / 32):.1f}" if steps else ""))
for op, n in by.most_common(28):
    print(f"  {op:10s} {n:.3e} {100 * n / tot:5.1f}%  lanes {byl[op] / max(1, n):5.1f}" + (f"  per warp-step {n / (steps / 32):6.1f}" if steps else ""))
# plateau segmentation: consecutive instructions whose executed counts are within 2 %
seg = []
for i, (n, t, s) in enumerate(data):
    if seg and abs(n - seg[-1][2]) <= 0.02 * max(n, seg[-1][2]) :
        seg[-1][1] = i; seg[-1][3] += n; seg[-1][4] += t
    else:
        seg.append([i, i, n, n, t])
print("segments (first..last instr, count per instr, total, share, lanes):")
for a, b, n, s, t in seg:
    if s > 0.004 * tot:
        print(f"  {a:5d}..{b:5d}  len {b - a + 1:4d}  exec/instr {n:.3e}  total {s:.3e}  {100 * s / tot:5.1f}%  lanes {t / max(1, s):5.1f}")
