"""Per-kernel census of the Blackwell-native instructions in the shipped library (no GPU needed):

    python tools/sass_census.py [path/to/libdfgan_b200.so] > profiles/rNN_sass_census.txt

Counts, per kernel of `cuobjdump -sass`, the SASS mnemonics that prove the tcgen05 / TMEM / TMA path
(/opt/skills/guides/B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA
tensor load / store, UBLKCP = bulk copy, LDGMC = multimem.ld_reduce (in-switch reduction) and system-scope / multicast
stores of the peer-memory exchange kernels, plus legacy HMMA (mma.sync; expected 0) and registers from the ELF header.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gan-codes_b200", "_C", "libdfgan_b200.so")
PAT = [("UTCMMA", re.compile(r"\bUTC[A-Z]*MMA\b")), ("LDTM", re.compile(r"\bLDTM\b")), ("STTM", re.compile(r"\bSTTM\b")),
       ("UTMALDG", re.compile(r"\bUTMALDG\b")), ("UTMASTG", re.compile(r"\bUTMASTG\b")),
       ("UBLKCP", re.compile(r"\bUBLKCP\b")), ("UTCBAR", re.compile(r"\bUTCBAR\b")), ("SYNCS", re.compile(r"\bSYNCS\b")),
       ("LDGMC", re.compile(r"\bLDGMC\b")), ("MC/SYS st", re.compile(r"\b(STG|ST)\.E\.[A-Z0-9.]*(MC|SYS)\b|\bREDG?\.[A-Z0-9.]*MC")),
       ("HMMA", re.compile(r"\bHMMA\b")), ("RED/ATOM", re.compile(r"\b(RED|REDG|ATOM|ATOMG|ATOMS)\b")), ("instr", re.compile(r"^\s+/\*[0-9a-f]{4}\*/"))]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    name = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and name:
            regs[name] = (int(m.group(1)), int(m.group(2)))
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for key, pat in PAT:
            if pat.search(line):
                counts[cur][key] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(counts)} kernels, sm_100a SASS (cuobjdump -sass)")
    hdr = [k for k, _ in PAT]
    print(f"{'kernel':78s} " + " ".join(f"{h:>8s}" for h in hdr) + "     regs  smem(static)")
    tot = collections.Counter()
    for (mangled, c), nice in zip(counts.items(), demangle):
        nice = re.sub(r"\(.*", "", nice.replace("dfgan::", "").replace("(anonymous namespace)::", ""))
        r = regs.get(mangled, ("", ""))
        print(f"{nice[:78]:78s} " + " ".join(f"{c[h]:8d}" for h in hdr) + f" {r[0]:>8} {r[1]:>8}")
        tot.update(c)
    print(f"{'TOTAL':78s} " + " ".join(f"{tot[h]:8d}" for h in hdr))


if __name__ == "__main__":
    main()
