"""Per-kernel counts of the Blackwell-only SASS instructions in libcnl_b200.so (cuobjdump -sass): python tools/sass_summary.py [out]

UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG / UTMASTG = TMA tensor loads / stores, LDTM = tcgen05.ld,
UTCBAR = tcgen05.commit, SYNCS = mbarrier try_wait / arrive.  Runs on a CPU-only box (no GPU needed)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "centernet-lightning_b200", "libcnl_b200.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMALDG.MULTICAST", "UTMASTG", "LDTM", "UTCBAR", "UTCBAR.MULTICAST", "SYNCS", "REDUX", "MATCH"]


def main() -> None:
    out = sys.argv[1] if len(sys.argv) > 1 else None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), stdout=subprocess.PIPE, text=True).stdout.split("\n")
    counts = collections.OrderedDict()
    fn = None
    it = iter(names)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = re.sub(r"\(.*", "", next(it)).replace("void ", "")
            counts[fn] = collections.Counter()
            counts[fn]["_instr"] = 0
            continue
        if fn is None or "/*" not in line or ";" not in line:
            continue
        counts[fn]["_instr"] += 1
        for k in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "SYNCS", "REDUX", "MATCH"):
            if re.search(r"\b" + k, line):
                counts[fn][k] += 1
                if k == "UTCHMMA" and ".2CTA" in line:
                    counts[fn]["UTCHMMA.2CTA"] += 1
                if k in ("UTMALDG", "UTCBAR") and "MULTICAST" in line:
                    counts[fn][k + ".MULTICAST"] += 1
    lines = ["# cuobjdump -sass centernet-lightning_b200/libcnl_b200.so (sm_100a), instruction counts per kernel",
             "# " + " ".join(f"{k:>9s}" for k in ["instr"] + KEYS) + "  kernel"]
    for fn, c in sorted(counts.items()):
        lines.append("  " + " ".join(f"{c[k]:9d}" for k in ["_instr"] + KEYS) + "  " + fn)
    text = "\n".join(lines) + "\n"
    if out:
        with open(out, "w") as f:
            f.write(text)
    else:
        sys.stdout.write(text)


if __name__ == "__main__":
    main()
