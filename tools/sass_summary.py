"""Per-kernel SASS instruction counts of libbasd_b200.so (cuobjdump -sass): the mnemonics that prove tcgen05 / TMEM / TMA use
(UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA load / store, UTCBAR = tcgen05.commit,
SYNCS = mbarrier) next to the legacy tensor-core path (HMMA) and packed fp32 (FFMA2).  usage: python tools/sass_summary.py > profiles/rN_sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vit_bias_aware_structural_distillation_b200", "libbasd_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "FFMA2", "FFMA", "MUFU", "SHFL", "BAR", "LDS", "STS", "LDG", "STG", "STAS"]


def strip_params(name):
    """Drops the trailing function parameter list '(...)' of a demangled name (template arguments stay)."""
    if not name.endswith(")"):
        return name
    depth = 0
    for i in range(len(name) - 1, -1, -1):
        if name[i] == ")":
            depth += 1
        elif name[i] == "(":
            depth -= 1
            if depth == 0:
                return name[:i]
    return name


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for k in KEYS:
                if op == k:
                    kernels[cur][k] += 1
    names = list(kernels)
    try:
        dm = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        demangle = dict(zip(names, dm))
    except Exception:
        pass
    print(f"# {os.path.relpath(LIB, ROOT)}: SASS instruction counts per kernel (sm_100a)")
    print("kernel," + ",".join(["total"] + KEYS))
    tot = collections.Counter()
    for n, c in kernels.items():
        short = strip_params(demangle.get(n, n)).replace("basd::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        short = re.sub(r"^void ", "", short).replace("(int)", "").replace("(bool)", "").replace(",", ";")
        print(short + "," + ",".join(str(c[k]) for k in ["_total"] + KEYS))
        tot.update(c)
    print("ALL," + ",".join(str(tot[k]) for k in ["_total"] + KEYS))


if __name__ == "__main__":
    sys.exit(main())
