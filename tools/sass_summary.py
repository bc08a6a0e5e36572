"""Per-kernel counts of the Blackwell-specific SASS mnemonics in libhmm_b200.so (cuobjdump -sass).  Run in the build container:

    python tools/sass_summary.py > profiles/<round>_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pytorch_hmm_b200", "lib", "libhmm_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FMNMX3", "HMMA",
         "REDUX", "MUFU.EX2", "DFMA", "BAR.SYNC", "BAR.ARV"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    cur, counts, sizes = None, collections.OrderedDict(), {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            sizes[cur] = 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            sizes[cur] += 1
            op = m.group(1)
            for w in WATCH:
                if op == w or op.startswith(w + ".") or (w == "MUFU.EX2" and op.startswith("MUFU.EX2")):
                    counts[cur][w] += 1
    print(f"# SASS mnemonic counts per kernel, {os.path.relpath(LIB, ROOT)} (sm_100a), from `cuobjdump -sass`")
    print("# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops")
    for fn, c in counts.items():
        if sizes[fn] == 0:
            continue
        name = demangle(fn)
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("hmmb200::", "")
        hits = ", ".join(f"{k}={v}" for k, v in c.items())
        print(f"{name:70s} {sizes[fn]:6d} instr   {hits}")


if __name__ == "__main__":
    sys.exit(main())
