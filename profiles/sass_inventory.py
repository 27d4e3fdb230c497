"""Regenerates r2_sass_inventory.txt: per-kernel SASS instruction counts of the built library (runs on CPU: cuobjdump + c++filt).

    python profiles/sass_inventory.py > profiles/r2_sass_inventory.txt
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "aukit_b200", "lib", "libaukit_cuda.so")
COLS = [("UBLKCP", r"^UBLKCP"), ("SYNCS", r"^SYNCS"), ("LDGSTS", r"^LDGSTS"), ("FFMA2", r"^FFMA2"), ("FMNMX3", r"^FMNMX3"),
        ("VIMNMX", r"^VIMNMX|^VIADDMNMX"), ("fp64", r"^DFMA|^DADD|^DMUL")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = [], None
    ins = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)")
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = {"name": m.group(1), "instr": 0, **{c: 0 for c, _ in COLS}}
            kernels.append(cur)
            continue
        m = ins.match(line)
        if m and cur is not None:
            op = m.group(1)
            cur["instr"] += 1
            for c, pat in COLS:
                if re.match(pat, op):
                    cur[c] += 1
    names = subprocess.run(["c++filt"], input="\n".join(k["name"] for k in kernels), capture_output=True, text=True, check=True).stdout.splitlines()
    print("SASS inventory of aukit_b200/lib/libaukit_cuda.so (cuobjdump -sass, sm_100a), final round-2 build (profiles/sass_inventory.py).")
    print("UBLKCP = bulk async copy (TMA, cp.async.bulk), SYNCS = mbarrier operations, LDGSTS = cp.async (16-byte global -> shared),")
    print("FFMA2 = packed f32x2 FMA (sm_100), FMNMX3 / VIMNMX* = 3-input float / integer min-max family, fp64 = DFMA + DADD + DMUL.")
    tensor = len(re.findall(r"UTMALDG|UTMASTG|UTCMMA|HMMA|IMMA|QMMA|OMMA", sass))
    print("UTMALDG / UTMASTG (tensor-map TMA) and tensor-core instructions in the whole library: %d -- nothing on this path is a contraction." % tensor)
    print()
    print("%-70s %6s" % ("kernel", "instr") + "".join(" %6s" % c for c, _ in COLS))
    tot = {"instr": 0, **{c: 0 for c, _ in COLS}}
    for k, n in zip(kernels, names):
        n = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", n)
        n = re.sub(r"\(.*$", "", n)
        print("%-70s %6d" % (n[:70], k["instr"]) + "".join(" %6d" % k[c] for c, _ in COLS))
        for key in tot:
            tot[key] += k[key]
    print("%-70s %6d" % ("total (%d kernels)" % len(kernels), tot["instr"]) + "".join(" %6d" % tot[c] for c, _ in COLS))


if __name__ == "__main__":
    sys.exit(main())
