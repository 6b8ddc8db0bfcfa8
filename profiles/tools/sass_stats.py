#!/usr/bin/env python
"""Static SASS census of one kernel in the built library: instruction mix by pipe-relevant mnemonic.
usage: sass_stats.py [lib.so] [substring of the mangled kernel name]"""
import collections, re, subprocess, sys

lib = sys.argv[1] if len(sys.argv) > 1 else "synthpy_b200/csrc/libsynthpy_b200.so"
pat = sys.argv[2] if len(sys.argv) > 2 else "k_propagateIdLi0ELb0ELb0E"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, funcs = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        funcs[cur].append(m.group(1))
for name, ins in funcs.items():
    if pat not in name:
        continue
    c = collections.Counter(i.split(".")[0] for i in ins)
    keys = ["DFMA", "DADD", "DMUL", "DSETP", "F2F", "LDG", "LDC", "LDS", "STL", "LDL", "MOV", "IMAD", "ISETP", "BRA", "SHFL", "MUFU"]
    print(name)
    print("  total", len(ins), " ".join(f"{k}={c[k]}" for k in keys if c[k]))
