"""Per-kernel counts of the tensor-core / TMA / TMEM SASS instructions in the built library (no GPU needed):
    python tools/sass_summary.py > profiles/r2/sass_tensor_instructions.txt
UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA bulk tensor load / store, LDTM = tcgen05.ld, HMMA = legacy mma.sync, SYNCS = mbarrier ops,
ACQBULK / griddepcontrol (programmatic dependent launch) shown as PDL."""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "efficientconformer_b200/libeffconf_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
print("# Tensor-core / TMA / TMEM instructions per kernel of libeffconf_b200.so (cuobjdump -sass, sm_100a), end of round 2 (tools/sass_summary.py)")
print("# UTCHMMA = tcgen05.mma (kind::f16 / tf32), UTMALDG / UTMASTG = TMA bulk tensor load / store, LDTM = tcgen05.ld (TMEM -> registers),")
print("# HMMA = legacy mma.sync, SYNCS = mbarrier operations.  Kernels without any of these are CUDA-core / bandwidth kernels and are omitted.\n")
cur, counts, order = None, collections.defaultdict(collections.Counter), []
it = iter(names)
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next(it).split("(")[0]
        order.append(cur)
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        for key in ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "HMMA", "SYNCS"):
            if op.startswith(key):
                counts[cur][key] += 1
for k in order:
    c = counts[k]
    if any(c[x] for x in ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "HMMA")):
        print(f"{k:<110} " + " ".join(f"{x}={c[x]}" for x in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "HMMA", "SYNCS") if c[x]))
