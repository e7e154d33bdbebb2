#!/usr/bin/env python
"""FP64 instruction mix of the main (largest backward-branch) loop of a kernel in the built library.

    python tools/sass_loop_count.py _Z11grid_kernelILi1ELi2ELi3EEv8GridArgs [...]

Used for the SASS counts quoted in DESIGN.md / profiles/ (cuobjdump -sass on the in-tree .so)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "eigensolver_b200", "libeigensolver_b200.so")


def functions():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, body = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
        elif cur and re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", line):
            body[cur].append(line)
    return body


def main_loop(lines):
    """the innermost backward-branch loop with more than 100 DFMA: the integration step loop"""
    addr = lambda l: int(re.search(r"/\*([0-9a-f]{4,})\*/", l).group(1), 16)
    loops = []
    for l in lines:
        if " BRA" in l:
            m = re.search(r"0x([0-9a-f]+)", l.split("BRA")[1])
            if m and int(m.group(1), 16) < addr(l):
                loops.append((addr(l) - int(m.group(1), 16), int(m.group(1), 16), addr(l)))
    best = None
    for span, lo, hi in sorted(loops):
        body = [l for l in lines if lo <= addr(l) <= hi]
        if sum(" DFMA" in l for l in body) > 100:
            best = body
            break
    return best or []


def mix(lines):
    c = collections.Counter()
    for l in lines:
        m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
        if m:
            c[m.group(1)] += 1
    return c


if __name__ == "__main__":
    body = functions()
    for name in sys.argv[1:]:
        hits = [f for f in body if name in f]
        for f in hits:
            loop = main_loop(body[f])
            c = mix(loop)
            fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DMUL", "DADD", "MUFU", "DSETP"))
            keys = ("DFMA", "DMUL", "DADD", "MUFU", "LDS", "LDL", "STL", "CALL")
            print("%s: loop %d instr, FP64-pipe %d  %s" % (f, len(loop), fp64, {k: c[k] for k in keys if c[k]}))
