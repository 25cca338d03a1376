"""SASS opcode histogram of libgaitk.so per kernel (static instruction counts): python scratch/sass_opcodes.py > profiles/rN_sass_opcodes.txt"""
import collections, re, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
lib = next(ROOT.glob("towards-*_b200/libgaitk.so"))
COLS = "UTCHMMA UTCQMMA LDTM STTM UTCBAR UTCATOMSWS UBLKCP UTMALDG LDGSTS SYNCS HMMA FFMA2 FFMA FMUL2 FADD2 DFMA MUFU F2FP SHFL BAR LDS STS LDG STG LDL STL".split()
sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kern = None; hist = collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m: kern = m.group(1); hist[kern] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern: hist[kern][m.group(1).split(".")[0]] += 1; hist[kern]["_total"] += 1
print("# SASS opcode histogram of libgaitk.so (cuobjdump -sass, sm_100a), per kernel; static instruction counts")
print("# columns: total | " + " ".join(COLS))
for k, h in sorted(hist.items(), key=lambda kv: -kv[1]["_total"]):
    cells = " ".join(f"{c}={h[c]}" for c in COLS if h[c])
    print(f"{h['_total']:7d} | {cells}  :: {demangle(k)}")
