"""Instruction-level evidence for profiles/: per kernel family of libclipcap_b200.so, how many tcgen05 / TMA / bulk-copy /
warp-MMA instructions its SASS holds (cuobjdump -sass), and registers / spills from the ptxas logs of the build.
    python scripts/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "clipcap_b200", "libclipcap_b200.so")
OPS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "HMMA", "LDSM"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    fam = collections.defaultdict(lambda: collections.Counter())
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            dem = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"^void ", "", dem)
            name = re.sub(r"cc::\(anonymous namespace\)::|cc::", "", name)
            cur = re.sub(r"[<(].*", "", name)
            fam[cur]["variants"] += 1
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for o in OPS:
                if op.startswith(o):
                    fam[cur][o] += 1
    print("# SASS evidence (cuobjdump -sass clipcap_b200/libclipcap_b200.so, sm_100a): instruction counts per kernel family")
    print("# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UTMAREDG = TMA tensor load/store/reduce-add,")
    print("# UBLKCP = cp.async.bulk, HMMA = mma.sync, LDSM = ldmatrix")
    print(f"{'kernel':36s}{'variants':>9s}" + "".join(f"{o:>9s}" for o in OPS))
    for k in sorted(fam, key=lambda k: -sum(fam[k][o] for o in OPS)):
        c = fam[k]
        if sum(c[o] for o in OPS) == 0:
            continue
        print(f"{k[:35]:36s}{c['variants']:9d}" + "".join(f"{c[o]:9d}" for o in OPS))
    print("\n# ptxas -v (registers / spill stores) of the kernels, from build/csrc/*.ptxas.log")
    for log in sorted(glob.glob(os.path.join(ROOT, "build", "csrc", "*.ptxas.log"))):
        txt = open(log).read()
        for m in re.finditer(r"Function properties for (\S+)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores.*\n.*Used (\d+) registers", txt):
            dem = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"cc::\(anonymous namespace\)::|cc::|^void ", "", dem)
            name = re.sub(r"\(.*", "", name)
            print(f"{name[:60]:62s} regs {int(m.group(4)):4d}  spill stores {int(m.group(3)):4d} B")


if __name__ == "__main__":
    sys.exit(main())
