"""Per-kernel SASS digest of the shipped library: instruction counts that show which hardware paths a kernel uses.
    python tools/sass_digest.py eqf_vio_b200/csrc/libeqvio_b200.so > profiles/r02_sass_digest.md
UTMALDG = TMA tile load (cp.async.bulk.tensor), SYNCS = mbarrier, DMMA = fp64 tensor core (mma.sync m8n8k4.f64),
UTC*MMA / LDTM = tcgen05 MMA / TMEM load, LDGSTS = cp.async, UBLKCP = cp.async.bulk, LDS / STS = shared-memory traffic, BAR = CTA barriers."""
import re
import subprocess
import sys
from collections import Counter

lib = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
names = re.findall(r"Function : (\S+)", out)
dm = dict(zip(names, demangle))
keys = ["UTMALDG", "UTMASTG", "SYNCS", "DMMA", "UTC", "LDTM", "STTM", "IMMA", "LDS", "STS", "LDGSTS", "UBLKCP", "LDG", "STG", "BAR", "ATOM", "RED", "MUFU", "DFMA", "DADD", "DMUL", "SHFL"]
print("| kernel | SASS instr | " + " | ".join(keys) + " |")
print("|---|---|" + "---|" * len(keys))
blocks = re.split(r"\n\s*Function : ", out)
for b in blocks[1:]:
    name = b.split("\n", 1)[0].strip()
    ins = re.findall(r"^\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", b, flags=re.M)
    c = Counter()
    for i in ins:
        base = i.split(".")[0]
        for k in keys:
            if base.startswith(k):
                c[k] += 1
                break
    pretty = dm.get(name, name).replace("eqvio::", "")
    pretty = pretty.replace("void ", "")
    depth = 0
    for i, ch in enumerate(pretty):   # cut the argument list: the first "(" outside the template brackets
        depth += ch == "<"
        depth -= ch == ">"
        if ch == "(" and depth == 0:
            pretty = pretty[:i]
            break
    print(f"| `{pretty}` | {len(ins)} | " + " | ".join(str(c[k]) for k in keys) + " |")
