"""Instruction mix per kernel of the built library (cuobjdump -sass): the mnemonics that prove which
hardware path a kernel uses - UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA tile load),
SYNCS (mbarrier), LDGSTS (cp.async), FFMA2 / FFMA (packed / scalar fp32 FMA), HMMA / IMMA (legacy
mma.sync: must stay 0).
usage: python tools/sass_mix.py [nas_3d_unet_b200/lib/libnas3d_b200.so] [name-filter] > profiles/sass_mix.txt"""
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "LDTM", "UTMALDG", "SYNCS", "LDGSTS", "FFMA2", "FFMA", "HMMA", "IMMA", "LDS", "STS", "LDG", "STG", "RED", "ATOM"]


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "nas_3d_unet_b200", "lib", "libnas3d_b200.so")
    flt = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    name, mix, names = None, {}, []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            names.append(name)
            mix[name] = Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and name:
            op = m.group(1)
            mix[name]["_total"] += 1
            key = next((k for k in KEYS if op == k or (k not in ("FFMA", "LDS", "STS", "LDG", "STG") and op.startswith(k))), op)
            mix[name][key] += 1
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    pretty = {n: re.sub(r"\(.*", "", d).replace("void ", "").replace("nas3d::", "") for n, d in zip(names, dem)}
    print("# %s" % os.path.relpath(path, ROOT))
    print("%-78s %7s " % ("kernel", "total") + " ".join("%7s" % k for k in KEYS))
    tot = Counter()
    for n in sorted(names, key=lambda n: pretty[n]):
        if flt in pretty[n]:
            c = mix[n]
            print("%-78s %7d " % (pretty[n][-78:], c["_total"]) + " ".join("%7d" % c[k] for k in KEYS))
            for k in KEYS:
                tot[k] += c[k]
    print("%-78s %7s " % ("ALL KERNELS", "") + " ".join("%7d" % tot[k] for k in KEYS))


if __name__ == "__main__":
    main()
