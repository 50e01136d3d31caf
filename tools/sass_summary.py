"""Per-kernel SASS opcode summary of the shipped library (runs where cuobjdump is installed; no GPU needed).
python tools/sass_summary.py > profiles/r2_sass_opcode_summary.txt"""
import collections
import os
import re
import subprocess
import sys

so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ocr-system_b200", "liblumina_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
kern, ops = None, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        ops[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        ops[kern][m.group(1).split(".")[0]] += 1
dem = subprocess.run(["cu++filt"] + list(ops), capture_output=True, text=True).stdout.splitlines() if ops else []
print(f"# SASS opcode summary of ocr-system_b200/liblumina_b200.so  (cuobjdump -sass; arch = {', '.join(arch)})")
print("# kernel | instructions | notable opcodes (count)")
NOTE = ("IDP", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "ST.ASYNC", "ATOMS", "ATOMG", "RED", "SHFL", "VOTE", "MATCH", "LDG", "STG",
        "LDS", "STS", "PRMT", "SHF", "IMAD", "DFMA", "DMUL", "DADD", "MUFU", "BAR", "UCGABAR", "CCTL", "MEMBAR", "ERRBAR", "LDSM",
        "HMMA", "IMMA", "F2I", "I2F", "POPC", "FLO", "LOP3", "LEA", "ISETP", "BRA", "NANOSLEEP", "ACQBULK", "FENCE")
tot = collections.Counter()
for (k, c), name in zip(ops.items(), dem or list(ops)):
    n = sum(c.values())
    tot.update(c)
    short = re.sub(r"\(.*", "", name).replace("lumina::", "").replace("void ", "")
    picks = [(o, c[o]) for o in NOTE if c.get(o)]
    print(f"{short} | {n} | " + " ".join(f"{o}:{v}" for o, v in picks))
print("# whole library: " + " ".join(f"{o}:{tot[o]}" for o in NOTE if tot.get(o)))
