"""Blackwell-specific SASS instructions per kernel of libngpb200.so:
    cuobjdump -sass blender-ngp_b200/libngpb200.so | python tools/sass_inventory.py > profiles/r02_sass_inventory.txt"""
import collections
import re
import subprocess
import sys

PAT = re.compile(r"\b(UTCHMMA|UTCQMMA|LDTM|STTM|UTCBAR|UBLKCP|UTMALDG|UTMASTG|SYNCS|UTCATOMSWS|REDG|ATOMG|LDGSTS|HMMA|LDG\.E\.64|LDG\.E\.128)\b")
fn, counts = None, collections.OrderedDict()
for line in sys.stdin:
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1); counts[fn] = collections.Counter(); continue
    if fn:
        for t in PAT.findall(line):
            counts[fn][t] += 1
print("# SASS inventory of blender-ngp_b200/libngpb200.so (cuobjdump -sass, sm_100a): Blackwell-specific instructions per kernel")
print("# UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit -> mbarrier, UBLKCP = cp.async.bulk (TMA 1-D bulk copy), SYNCS = mbarrier ops,")
print("# UTCATOMSWS = tcgen05.alloc / dealloc, REDG / ATOMG = global reductions / atomics, LDG.E.64 / .128 = vector loads\n")
tot = collections.Counter()
for f, c in counts.items():
    if not c:
        continue
    dem = re.sub(r"\(.*", "", subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip())
    print(f"{dem:70s} " + "  ".join(f"{k}={v}" for k, v in sorted(c.items())))
    tot.update(c)
print("\nTOTAL " + "  ".join(f"{k}={v}" for k, v in sorted(tot.items())))
