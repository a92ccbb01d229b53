"""Instruction mix per kernel of a CUDA object / library (cuobjdump -sass): FFMA2 / FFMA / LDS / MOV
counts, used to check that the packed-FMA conversion really lands as FFMA2 with no extra moves.
usage: python tools/sass_mix.py file.o [name-filter]"""
import re
import subprocess
import sys
from collections import Counter


def main():
    path = sys.argv[1]
    flt = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    name = None
    mix = {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name)
            mix[name] = Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and name:
            mix[name][m.group(1)] += 1
    keys = ["FFMA2", "FFMA", "FMUL", "FADD", "LDS", "MOV", "IMAD", "IADD3", "LEA", "STS", "LDG", "STG"]
    print("%-70s %6s " % ("kernel", "total") + " ".join("%6s" % k for k in keys))
    for n, c in mix.items():
        if flt in n:
            print("%-70s %6d " % (n[-70:], sum(c.values())) + " ".join("%6d" % c[k] for k in keys))


if __name__ == "__main__":
    main()
