"""Per-kernel SASS opcode evidence (run after the build): which kernels issue FP64 tensor-core MMAs (DMMA), asynchronous
global->shared copies (LDGSTS = cp.async), double-precision FMAs, 16-byte global stores, plus registers / spills from
ptxas -v.      python tools/sass_counts.py > profiles/r02_sass_opcodes.txt"""
import os
import re
import subprocess
from collections import Counter

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "stopro_b200", "csrc")
OPS = ("DMMA", "DFMA", "DMUL", "DADD", "LDGSTS", "LDS", "LDG", "STG", "BAR", "MUFU")


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().replace("pigp::", "")
    except OSError:
        return name


print("# cuobjdump -sass of the objects of libpigp.so (sm_100a): opcode counts per kernel")
print("# (DMMA = mma.sync.m8n8k4.f64, the FP64 tensor-core instruction of sm_100a -- tcgen05 has no f64 kind; LDGSTS = cp.async;")
print("#  STG.128 = 16-byte global stores)")
for obj in ("pigp_dense.o", "pigp_assemble.o", "pigp_matern.o", "pigp_dist.o", "pigp_capi.o"):
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(CSRC, obj)], capture_output=True, text=True).stdout
    fn, cnt, wide, tot = None, Counter(), 0, 0

    def flush():
        if fn:
            body = " ".join(f"{o}={cnt[o]}" for o in OPS if cnt[o])
            print(f"{obj:16s} {demangle(fn)[:70]:70s} total={tot:5d} {body} STG.128={wide}")

    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            flush()
            fn, cnt, wide, tot = m.group(1), Counter(), 0, 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m:
            tot += 1
            cnt[m.group(1)] += 1
            if m.group(1) == "STG" and ".128" in m.group(2):
                wide += 1
    flush()
print()
print("# ptxas -v: registers and spills of every kernel")
for log in ("pigp_dense.ptxas.log", "pigp_assemble.ptxas.log", "pigp_matern.ptxas.log", "pigp_dist.ptxas.log"):
    path = os.path.join(CSRC, log)
    if not os.path.exists(path):
        continue
    cur = None
    for line in open(path):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = demangle(m.group(1))
        m2 = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m2 and cur:
            spill = (m2.group(1), m2.group(2))
        m3 = re.search(r"Used (\d+) registers", line)
        if m3 and cur:
            print(f"{cur[:80]:80s} registers={m3.group(1):4s} spill_st/ld={spill[0]}/{spill[1]} B")
            cur = None
