"""Per-kernel list of the Blackwell-specific SASS mnemonics in the built library (profiles/r2_sass_mnemonics.txt).
    python tools/sass_mnemonics.py > profiles/r2_sass_mnemonics.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "raft_optical_flow_b200", "libraftcorr_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
pat = re.compile(r"^(UTC|UTMA|UBLKCP|LDTM|STTM|SYNCS|ELECT|REDUX|ACQBULK|LDGSTS|REDG|FFMA2|FMUL2)")
print("# cuobjdump -sass raft_optical_flow_b200/libraftcorr_b200.so: tcgen05 / TMA / TMEM / bulk-copy / mbarrier mnemonics per kernel")
print("# (UTCHMMA = tcgen05.mma kind::f16, UTCQMMA = kind::f8f6f4, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = cp.async.bulk.tensor load/store,")
print("#  UBLKCP = cp.async.bulk (linear), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops)")
fn, counts = None, None
def flush():
    if fn is not None:
        print(fn)
        if counts:
            print("    " + ", ".join(f"{k} x{v}" for k, v in sorted(counts.items())))
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        fn, counts = m.group(1), collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
    if m and fn and pat.match(m.group(1)):
        counts[m.group(1)] += 1
flush()
