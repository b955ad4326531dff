"""Opcode histogram of every kernel in lib/libvlb200.so (cuobjdump -sass), written to profiles/.

    python tools/sass_histogram.py [out.txt]

The lines that matter for the tensor-core claim: UTCHMMA (tcgen05.mma), UTMALDG / UTMASTG (TMA loads / stores, incl.
.IM2COL), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), SYNCS (mbarrier), and for the HBM-bound kernels MUFU / FFMA /
HFMA2 / LDG.E.128 / STG.E.128 counts."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "video-learning-tf_b200", "lib", "libvlb200.so")
KEY = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "ELECT",
       "MUFU", "FFMA", "FFMA2", "HFMA2", "HMNMX2", "LDG", "STG", "LDS", "STS", "RED", "ATOM", "SHFL", "BAR", "F2FP")


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_histogram.txt")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            funcs[cur][m.group(1)] += 1
            funcs[cur]["__total__"] += 1
            full = m.group(1) + m.group(2)
            for tag in (".IM2COL", ".128", ".2CTA"):
                if tag in full and m.group(1) in ("UTMALDG", "UTMASTG", "LDG", "STG", "LDS", "STS", "UTCHMMA"):
                    funcs[cur][m.group(1) + tag] += 1
    demangle = subprocess.run(["cu++filt"] + list(funcs), capture_output=True, text=True)
    names = demangle.stdout.splitlines() if demangle.returncode == 0 else list(funcs)
    with open(out_path, "w") as fh:
        fh.write("# cuobjdump -sass %s : opcode counts per kernel (static instruction counts)\n" % os.path.relpath(LIB, ROOT))
        tot = collections.Counter()
        for (mangled, cnt), name in zip(funcs.items(), names):
            short = re.sub(r"\((?!anonymous).*", "", name.replace("(anonymous namespace)::", "").replace("<unnamed>::", ""))
            items = ["%s=%d" % (k, cnt[k]) for k in sorted(cnt) if k != "__total__" and (k.split(".")[0] in KEY)]
            fh.write("%-100s total=%5d  %s\n" % (short[:100], cnt["__total__"], " ".join(items)))
            tot.update(cnt)
        fh.write("\n# library totals: " + " ".join("%s=%d" % (k, tot[k]) for k in sorted(tot)
                                                  if k.split(".")[0] in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR",
                                                                         "SYNCS", "MUFU")) + "\n")
    print(out_path)


if __name__ == "__main__":
    main()
