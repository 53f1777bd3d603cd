#!/usr/bin/env python
"""Static checks of the built library's SASS (cuobjdump works without a GPU) and the evidence file the docs cite:
which kernels use the TMA path (UTMALDG + mbarrier SYNCS), that the ingest gathers are 256-bit, that the strip egress
prefetches into L1, and that the Phase egress clips NaN to 1.0 with an explicit select (OpenCV's max(min(v,1),0)) instead
of a .SAT folded into the producing FFMA (round-1 hardware failure).  Usage: python tools/check_sass.py [out.txt]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "live-video-magnification_b200", "libmagcore_b200.so")


def kernels():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    cur, body = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"mc::\(anonymous namespace\)::", "", cur)
            body[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            body[cur].append(re.sub(r"\s*/\*.*?\*/\s*", " ", line).strip())
    return body


def main():
    body = kernels()
    count = lambda k, pat: sum(1 for l in body[k] if re.search(pat, l))
    lines, ok = [], True
    lines.append("SASS evidence of libmagcore_b200.so (cuobjdump -sass; tools/check_sass.py)\n")
    lines.append(f"{'kernel':90s} instr  UTMALDG SYNCS LDG.256 CCTL.PF1 SHFL  BAR")
    for k in body:
        lines.append(f"{k[:90]:90s} {len(body[k]):5d}  {count(k, 'UTMALDG'):7d} {count(k, 'SYNCS'):5d} {count(k, r'LDG\.E\.\S*256'):7d} "
                     f"{count(k, r'CCTL\.E\.PF1'):8d} {count(k, 'SHFL'):4d} {count(k, r'BAR\.SYNC'):4d}")

    def need(cond, what):
        nonlocal ok
        lines.append(("ok   " if cond else "FAIL ") + what)
        ok = ok and cond

    lines.append("")
    tma = [k for k in body if k.startswith("void k_level<0, true")]
    need(bool(tma) and all(count(k, "UTMALDG") >= 1 and count(k, "SYNCS") >= 2 for k in tma), "k_level<f32, TMA, *>: cp.async.bulk.tensor (UTMALDG) + mbarrier (SYNCS)")
    pre = [k for k in tma if k.startswith("void k_level<0, true, true")]
    need(bool(pre) and all(count(k, "UTMALDG") == 3 for k in pre), "k_level<f32, TMA, PREFETCH>: three bulk-tensor copies (input window + both state tiles)")
    r9 = [k for k in body if k.startswith("k_riesz_analysis(") or k.startswith("k_riesz_collapse(")]
    need(len(r9) == 2 and all(count(k, "UTMALDG") == 1 and count(k, "SYNCS") >= 2 for k in r9), "k_riesz_analysis / k_riesz_collapse: 9x9 input tile by one bulk-tensor copy")
    ing = [k for k in body if k.startswith("void k_ingest_lab<")]
    need(bool(ing) and all(count(k, r"LDG\.E\.\S*256") >= 8 for k in ing), "k_ingest_lab: 256-bit LUT gathers (LDG.E.*.256)")
    strip = [k for k in body if k.startswith("void k_egress_strip<3")]
    need(bool(strip) and all(count(k, r"CCTL\.E\.PF1") >= 8 and count(k, r"BAR\.SYNC") == 0 for k in strip), "k_egress_strip<3>: L1 prefetches, no barrier")
    rz = [k for k in body if k.startswith("k_riesz_egress") or "k_riesz_egress(" in k]
    need(bool(rz) and all(count(k, r"FSETP\.NAN") >= 3 and count(k, "FSEL") >= 3 and count(k, r"FFMA\.SAT") == 0 for k in rz),
         "k_riesz_egress: explicit NaN -> 1.0 select before the gamma (FSETP.NAN + FSEL), no FFMA.SAT")
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text)
    print(text)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
