#!/usr/bin/env python
"""Count the Blackwell-path SASS mnemonics of the built library per kernel (`cuobjdump -sass`, runs without a GPU) and
write profiles/<tag>_sass_evidence.md.  usage: python scripts/sass_evidence.py <tag>"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "SYNCS", "LDGMC", "REDG", "ATOMG", "FFMA", "HMMA"]


def main():
    tag = sys.argv[1]
    so = os.path.join(ROOT, "ugaitnet_b200", "libugaitnet_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            if op in OPS:
                per[cur][op] += 1
    names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    with open(os.path.join(ROOT, "profiles", f"{tag}_sass_evidence.md"), "w") as f:
        f.write(f"# {tag}: SASS mnemonics of the built `libugaitnet_b200.so` (`cuobjdump -sass`, sm_100a)\n\n"
                "`UTCHMMA` = tcgen05.mma (kind::f16), `UTMALDG` = TMA tensor loads, `LDTM` = tcgen05.ld (TMEM -> registers),\n"
                "`UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier ops, `LDGMC` = multimem.ld_reduce (NVSwitch in-fabric reduction),\n"
                "`REDG` = red.global.add, `HMMA` = legacy mma.sync (must stay 0).\n\n")
        f.write("whole library: " + ", ".join(f"{o} {tot[o]}" for o in OPS) + f"; {len(per)} kernels\n\n")
        f.write("| kernel | " + " | ".join(OPS) + " |\n|---|" + "---:|" * len(OPS) + "\n")
        for (k, c), n in zip(per.items(), names):
            if not any(c[o] for o in ("UTCHMMA", "UTMALDG", "LDTM", "LDGMC")):
                continue
            short = re.sub(r"\(.*", "", n).replace("void ", "")
            f.write(f"| `{short}` | " + " | ".join(str(c[o]) for o in OPS) + " |\n")
    print("whole library:", dict(tot), len(per), "kernels")


if __name__ == "__main__":
    main()
