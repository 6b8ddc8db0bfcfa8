#!/usr/bin/env python
"""bench.parity_check on a benchmarked configuration, one JSON line per (n, offset) window.
usage: parity_report.py WORKLOAD N OFFSET [OFFSET ...]"""
import json, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import bench
from synthpy_b200 import beam as B, domain as Dm

w, n, offs = sys.argv[1], int(sys.argv[2]), [int(v) for v in sys.argv[3:]]
a = bench.parse(["--workload", w])
ne = bench.build_ne(a, "cuda")
dom = Dm.ScalarDomain(bench.LENGTHS, a.grid)
dom.external_ne(ne)
odom = bench.cpu_setup(ne.cpu().numpy(), a)
del ne
beam = B.Beam(int(a.rays), bench.BEAM_R, bench.BEAM_DIV, bench.EXTENT, device=True, seed=2, beam_type="circular")
for off in offs:
    print(json.dumps(dict(workload=w, **bench.parity_check(a, dom, beam, odom, n, ray_offset=off))), flush=True)
