"""Per-kernel counts of the Blackwell-only SASS mnemonics in the built libqst.so (proof that the hot
kernels are tcgen05 / TMEM / TMA code and not recompiled mma.sync):

    python profiles/sass_excerpt.py > profiles/r02_sass_tcgen05.txt

UTCHMMA = tcgen05.mma (kind::f16), LDTM / STTM = tcgen05.ld / .st, UTMALDG = cp.async.bulk.tensor,
UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, UTCATOMSWS / UTCALLOC-like = TMEM alloc."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "quadruplet-sentence-transformer_b200", "libqst.so")
PAT = re.compile(r"\b(UTCHMMA[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTMALDG[.\w]*|UTCBAR[.\w]*|UTCCP[.\w]*|SYNCS[.\w]*|HMMA[.\w]*)")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = PAT.search(line)
        if m:
            per[cur][m.group(1)] += 1
    print("# cuobjdump -sass libqst.so: tcgen05 / TMEM / TMA / mbarrier mnemonics per kernel (count)")
    for k, c in per.items():
        if not c:
            continue
        print(k)
        for name, n in sorted(c.items()):
            print(f"    {name:40s} {n}")


if __name__ == "__main__":
    sys.exit(main())
