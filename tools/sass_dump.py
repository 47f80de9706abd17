#!/usr/bin/env python
"""Dump a SASS function with decoded control fields (stall, yield, wbar, rbar, wait mask, reuse)."""
import re
import subprocess
import sys


def functions(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    res = {}
    for b in out.split("Function : ")[1:]:
        name, body = b.split("\n", 1)
        res[name.strip()] = body
    return res


PAIR = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/\s*\n\s+/\* (0x[0-9a-f]{16}) \*/")


def decode(body):
    ins = []
    for m in PAIR.finditer(body):
        addr = int(m.group(1), 16)
        lo, hi = int(m.group(3), 16), int(m.group(4), 16)
        w = (hi << 64) | lo
        ctl = dict(stall=(w >> 105) & 0xF, yld=(w >> 109) & 1, wbar=(w >> 110) & 7, rbar=(w >> 113) & 7,
                   wait=(w >> 116) & 0x3F, reuse=(w >> 122) & 0xF)
        ins.append((addr, m.group(2).strip(), w, ctl))
    return ins


if __name__ == "__main__":
    fn = functions(sys.argv[1])
    name = [n for n in fn if sys.argv[2] in n][0]
    lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
    hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
    print(name)
    for addr, text, w, c in decode(fn[name]):
        if lo <= addr <= hi:
            print(f"{addr:05x} {text:52s} st={c['stall']:2d} y={c['yld']} wb={c['wbar']} rb={c['rbar']} "
                  f"wt={c['wait']:06b} ru={c['reuse']:04b}  {w:032x}")
