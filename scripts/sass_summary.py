"""SASS mnemonic counts per kernel of a built library (the evidence that the hot kernels really use tcgen05 / TMEM / TMA /
cluster barriers / programmatic dependent launch): python scripts/sass_summary.py deepflows_b200/lib/libdfb200.so > profiles/<round>_sass_summary.txt"""
import collections
import re
import subprocess
import sys

WATCH = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "SYNCS", "ELECT", "UCGABAR_ARV", "UCGABAR_WAIT", "ACQBULK", "HMMA", "FFMA",
         "LDG", "STG", "LDS", "STS", "ATOMG", "RED", "MEMBAR"]


def main(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    rows, name, counts, n = [], None, None, 0
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                rows.append((name, counts, n))
            name, counts, n = m.group(1), collections.Counter(), 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and name:
            n += 1
            counts[m.group(1).split(".")[0]] += 1
    if name:
        rows.append((name, counts, n))
    names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
    print("SASS mnemonic counts per kernel of %s (cuobjdump -sass, sm_100a)." % path.split("/")[-1])
    print("tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, tcgen05.commit -> UTCBAR, TMA -> UTMALDG, cluster barrier -> UCGABAR_ARV / UCGABAR_WAIT,")
    print("mbarrier -> SYNCS, elect.sync -> ELECT, griddepcontrol.wait -> ACQBULK (B200_PROFILING.md). HMMA would be legacy mma.sync.")
    total = collections.Counter()
    print()
    for (_, counts, n), d in sorted(zip(rows, names), key=lambda r: r[1]):
        total.update(counts)
        print(d)
        print("    instructions %d | %s" % (n, "  ".join("%s=%d" % (k, counts[k]) for k in WATCH if counts[k])))
    print()
    print("whole library: %d kernels | %s" % (len(rows), "  ".join("%s=%d" % (k, total[k]) for k in WATCH)))


if __name__ == "__main__":
    main(sys.argv[1])
