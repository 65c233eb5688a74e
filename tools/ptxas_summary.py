#!/usr/bin/env python
"""Per-kernel registers / spills / shared memory from the ptxas -v logs of the last build (csrc/build/*.ptxas.log)."""
import glob, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pat = sys.argv[1] if len(sys.argv) > 1 else ""
for log in sorted(glob.glob(os.path.join(root, "cuda-path-tracer-denoising_b200", "csrc", "build", "*.ptxas.log"))):
    lines = open(log).read().splitlines()
    for i, l in enumerate(lines):
        m = re.search(r"Compiling entry function '(\S+)'", l)
        if not m:
            continue
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0]
        if pat and pat not in name:
            continue
        blob = " ".join(lines[i + 1:i + 5])
        regs = re.search(r"Used (\d+) registers", blob); sp = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", blob)
        st = re.search(r"(\d+) bytes stack frame", blob); sm = re.search(r"(\d+) bytes smem", blob)
        print("%-70s regs %3s  stack %4s  spill st/ld %s/%s  smem %s" % (name[:70], regs.group(1) if regs else "?", st.group(1) if st else "?",
                                                                      sp.group(1) if sp else "?", sp.group(2) if sp else "?", sm.group(1) if sm else "0"))
