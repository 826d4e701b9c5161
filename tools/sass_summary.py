#!/usr/bin/env python
"""Per-kernel SASS evidence of the tensor-core / TMA paths (no GPU needed):

    python tools/sass_summary.py > profiles/sass_summary_r2.txt

Disassembles lct-gan_b200/lctgan/liblctgan_sm100.so with cuobjdump and counts, per kernel, the mnemonics that prove which
hardware path it uses (B200_PROFILING.md): tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, tcgen05.commit -> UTCBAR,
tcgen05.alloc -> UTCATOMSWS, TMA -> UTMALDG/UTMASTG/UBLKCP, legacy mma.sync -> HMMA, cp.async -> LDGSTS."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "lct-gan_b200", "lctgan", "liblctgan_sm100.so")
PATTERNS = ["UTC[A-Z]*MMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "LDGSTS",
            "SYNCS", "FFMA2?", "MUFU"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        mm = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not mm:
            continue
        op = mm.group(1).split(".")[0]
        counts[cur]["_total"] += 1
        for p in PATTERNS:
            if re.fullmatch(p, op):
                counts[cur][p] += 1
    names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print(f"# SASS mnemonic counts per kernel of {os.path.relpath(SO, ROOT)} (cuobjdump -sass; sm_100a)")
    print(f"# {'instr':>6} {'UTC*MMA':>7} {'LDTM':>5} {'UTCBAR':>6} {'TMEMalloc':>9} {'UTMALDG':>7} {'HMMA':>5} {'LDGSTS':>6}  kernel")
    for (mangled, c), name in zip(counts.items(), names):
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name)[:100]
        if not any(c[p] for p in ("UTC[A-Z]*MMA", "LDTM", "UTMALDG", "HMMA", "LDGSTS")) and "--all" not in sys.argv:
            continue
        print(f"  {c['_total']:>6} {c['UTC[A-Z]*MMA']:>7} {c['LDTM']:>5} {c['UTCBAR']:>6} {c['UTCATOMSWS']:>9} {c['UTMALDG']:>7} "
              f"{c['HMMA']:>5} {c['LDGSTS']:>6}  {name}")


if __name__ == "__main__":
    main()
