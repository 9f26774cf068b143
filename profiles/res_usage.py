"""Per-kernel resource usage of the built libqst.so (registers, stack = spills, static shared memory):

    python profiles/res_usage.py > profiles/r02_res_usage.txt

`cuobjdump -res-usage` on the sm_100a cubin; template instances of one kernel are folded into one line
(min..max).  STACK > 0 would mean local-memory spills or a local array."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "quadruplet-sentence-transformer_b200", "libqst.so")


def span(vals):
    return str(min(vals)) if min(vals) == max(vals) else f"{min(vals)}..{max(vals)}"


def main():
    out = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    names = re.findall(r"Function (\S+):", out)
    demangled = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    usage = re.findall(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", out)
    assert len(names) == len(usage) == len(demangled)
    per = collections.OrderedDict()
    for name, (reg, stack, shared, local) in zip(demangled, usage):
        base = re.sub(r"^void ", "", name)
        base = re.sub(r"[<(].*", "", base)
        per.setdefault(base, []).append((int(reg), int(stack), int(shared), int(local)))
    print("# cuobjdump -res-usage libqst.so (sm_100a); template instances folded, min..max")
    print(f"{'kernel':48s} {'instances':>9s} {'registers':>10s} {'stack B':>8s} {'static smem B':>14s} {'local B':>8s}")
    for base, rows in per.items():
        cols = list(zip(*rows))
        print(f"{base:48s} {len(rows):9d} {span(cols[0]):>10s} {span(cols[1]):>8s} {span(cols[2]):>14s} {span(cols[3]):>8s}")


if __name__ == "__main__":
    main()
