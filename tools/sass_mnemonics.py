#!/usr/bin/env python
"""cuobjdump -sass libsrnn_b200.so | python tools/sass_mnemonics.py > profiles/rN_sass_mnemonics.txt
Static counts of the tensor-core / TMEM / TMA / cluster mnemonics per kernel."""
import collections
import re
import subprocess
import sys

PAT = re.compile(r'\b(UTCHMMA(?:\.2CTA)?|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|UTCBAR(?:\.2CTA)?(?:\.MULTICAST)?|UCGABAR_ARV|ACQBULK|HMMA)\b')


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
    except Exception:
        return n


def main():
    cur, cnt = None, collections.OrderedDict()
    for line in sys.stdin:
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            cnt[cur] = collections.Counter()
            continue
        if cur is not None:
            for mm in PAT.finditer(line):
                cnt[cur][mm.group(1)] += 1
    print("# cuobjdump -sass libsrnn_b200.so (sm_100a): tensor-core / TMEM / TMA / cluster mnemonics per kernel (static counts)")
    print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / tcgen05.st (weights into tensor memory),")
    print("# UTMALDG / UTMASTG = cp.async.bulk.tensor load / store, UBLKCP = cp.async.bulk (shared::cta -> shared::cluster push of the")
    print("# partial logits), UTCBAR = tcgen05.commit, UCGABAR_ARV = cluster barrier, ACQBULK = griddepcontrol.wait (programmatic")
    print("# dependent launch).  HMMA would be mma.sync / wmma.")
    rows = sorted((demangle(k), ", ".join("%s x%d" % (a, b) for a, b in sorted(c.items()))) for k, c in cnt.items() if c)
    for n, v in rows:
        print("%-58s %s" % (n[:58], v))
    print("# kernels in the library: %d; with tcgen05.mma: %d; with HMMA: %d" % (
        len(cnt), sum(1 for c in cnt.values() if any(k.startswith("UTCHMMA") for k in c)), sum(1 for c in cnt.values() if c.get("HMMA"))))


if __name__ == "__main__":
    main()
