"""Per-kernel counts of the SASS mnemonics that prove the Blackwell features in the shipped library
(cuobjdump -sass epilogos_b200/libepilogos_b200.so): tcgen05.mma (UTC*MMA), tcgen05.ld (LDTM), TMA tensor loads (UTMALDG),
bulk copies (UBLKCP), mbarrier phase checks (SYNCS), plus the register count of every kernel.
    python tools/sass_summary.py > profiles/r02_sass_summary.txt          (runs without a GPU)"""
import re
import subprocess
from collections import Counter, OrderedDict
from pathlib import Path

lib = Path(__file__).resolve().parent.parent / "epilogos_b200" / "libepilogos_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", str(lib)], capture_output=True, text=True).stdout
regs = {}
name = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        name = m.group(1)
    m = re.search(r"REG:(\d+)", line)
    if m and name:
        regs[name] = int(m.group(1))
WANT = OrderedDict([("UTCIMMA", "tcgen05.mma kind::i8"), ("UTCHMMA", "tcgen05.mma kind::f16"), ("LDTM", "tcgen05.ld"),
                    ("UTMALDG", "TMA tensor load"), ("UBLKCP", "bulk copy"), ("UTCBAR", "tcgen05.commit"),
                    ("SYNCS", "mbarrier"), ("DFMA", "fp64 fma"), ("IDP", "dp4a")])
kern = None
counts = {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        counts[kern]["_total"] += 1
        for w in WANT:
            if op.startswith(w):
                counts[kern][w] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass of %s (sm_100a); static instruction counts per kernel" % lib.name)
print("# " + ", ".join("%s = %s" % kv for kv in WANT.items()))
print("%-8s %-5s %s  kernel" % ("instr", "regs", " ".join("%8s" % w for w in WANT)))
for (k, c), d in zip(counts.items(), demangle):
    short = re.sub(r"\(.*", "", d).replace("epi::", "").replace("void ", "")
    print("%-8d %-5s %s  %s" % (c["_total"], regs.get(k, ""), " ".join("%8d" % c[w] for w in WANT), short[:110]))
