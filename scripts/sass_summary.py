"""Per-kernel counts of the SASS mnemonics that show what a kernel is built on (B200_PROFILING.md, "What proves a
Blackwell-native kernel"): tcgen05 MMAs (UTC*MMA), tensor-memory loads / stores (LDTM / STTM), TMA (UTMALDG / UTMASTG / UBLKCP),
legacy tensor-core paths (HMMA / HGMMA: must be absent), cp.async (LDGSTS), system-scope polls of the peer exchange
(LDG.E.64.STRONG.SYS) and 16-byte global stores.

    python scripts/sass_summary.py [path/to/libdrsa_b200.so] > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MNEMONICS = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "HMMA", "HGMMA", "LDGSTS",
             "LDG.E.64.STRONG.SYS", "STG.E.128", "SYNCS", "MEMBAR")


# one instruction per SASS line: the first mnemonic of the table that matches (longest alternatives first)
PATTERN = re.compile(r"(?<![\w.])(" + "|".join(re.escape(m) for m in sorted(MNEMONICS, key=len, reverse=True)) + r")(?![\w])")


def kernel_name(full: str) -> str:
    """Demangled name without namespaces and argument list: drsa_tc_step_kernel<256, 0, 1>."""
    full = re.sub(r"^void\s+", "", full)
    full = full.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    full = re.sub(r"\((?:int|bool|unsigned int|long)\)", "", full)          # template arguments are printed as casts
    head = full.split("(")[0]
    base, _, targs = head.partition("<")
    return base.split("::")[-1] + ("<" + targs if targs else "")


def summarise(so_path: str):
    sass = subprocess.run(["cuobjdump", "-sass", so_path], capture_output=True, text=True, check=True).stdout
    mangled = [line.split("Function :")[1].strip() for line in sass.splitlines() if "Function :" in line]
    try:          # one call for all names
        plain = subprocess.run(["cu++filt"] + mangled, capture_output=True, text=True, check=True).stdout.splitlines()
    except (OSError, subprocess.CalledProcessError):
        plain = mangled
    names = dict(zip(mangled, plain if len(plain) == len(mangled) else mangled))
    out, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        if "Function :" in line:
            cur = kernel_name(names[line.split("Function :")[1].strip()])
            out[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        m = PATTERN.search(line)
        if m:
            out[cur][m.group(1)] += 1
    return out


if __name__ == "__main__":
    so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "drsa_audio_b200", "libdrsa_b200.so")
    table = summarise(so)
    print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)}: mnemonic counts per kernel (kernels with none of them omitted)")
    for k, c in table.items():
        if c:
            print(f"{k}: " + ", ".join(f"{mn} {n}" for mn, n in sorted(c.items())))
