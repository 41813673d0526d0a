"""Counts of Blackwell-native SASS mnemonics per kernel of libwiski_b200.so -> profiles/rNN_sass_summary.txt.
usage: python tools/sass_summary.py [out]   (needs cuobjdump, no GPU)"""
import re
import subprocess
import sys

PATS = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "SYNCS", "LDGSTS", "HMMA", "FFMA"]


def main(out="profiles/r02_sass_summary.txt", lib="online_gp_b200/csrc/libwiski_b200.so"):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    lines = ["# SASS evidence, %s (sm_100a): counts of mnemonics per kernel" % lib,
             "# UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, LDTM / STTM = tcgen05.ld / tcgen05.st,",
             "# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, LDGSTS = cp.async, HMMA = legacy mma.sync, FFMA = fp32 FMA", "",
             "%-100s " % "kernel" + " ".join("%8s" % p for p in PATS)]
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        cnt = {p: len(re.findall(r"\b%s" % p, f)) for p in PATS}
        if not any(cnt[p] for p in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "LDGSTS")) and cnt["FFMA"] < 256:
            continue
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().replace("CUtensorMap_st", "TMap")
        lines.append("%-100s " % dem[:98] + " ".join("%8d" % cnt[p] for p in PATS))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(*sys.argv[1:])
