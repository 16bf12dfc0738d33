"""tcgen05 / TMA mnemonics per kernel of the product library: cuobjdump -sass bin_b200/libbin_b200.so | python tools/sass_summary.py"""
import re
import sys

txt = sys.stdin.read()
out = ["# r02: cuobjdump -sass bin_b200/libbin_b200.so -- tcgen05 / TMA mnemonics per kernel (product library)", ""]
cur, counts = None, {}
MN = ("UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "UTCBAR")
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = {}
        continue
    if cur is None:
        continue
    for mn in MN:
        if re.search(r"\s" + re.escape(mn) + r"[\s.]", line):
            key = mn
            if mn == "UTCHMMA" and "UTCHMMA.2CTA" in line:
                key = "UTCHMMA.2CTA"
            counts[cur][key] = counts[cur].get(key, 0) + 1
            break
out.append("| kernel | " + " | ".join(MN) + " |")
out.append("|---|" + "---|" * len(MN))
for fn, c in counts.items():
    if any(c.get(k) for k in MN):
        out.append(f"| {fn} | " + " | ".join(str(c.get(k, 0)) for k in MN) + " |")
out.append("")
out.append(f"whole library: UTCHMMA.2CTA x{sum(c.get('UTCHMMA.2CTA', 0) for c in counts.values())}, "
           f"UTCHMMA (one CTA) x{sum(c.get('UTCHMMA', 0) for c in counts.values())}, "
           f"UTMALDG x{sum(c.get('UTMALDG', 0) for c in counts.values())}, LDTM x{sum(c.get('LDTM', 0) for c in counts.values())}")
out.append("")
out.append("excerpt (rdb_tail_pair_kernel):")
inpair = False
n = 0
for line in txt.splitlines():
    if "Function :" in line:
        inpair = "rdb_tail_pair" in line
    if inpair and ("UTCHMMA.2CTA" in line or "UTCBAR" in line or "UTMALDG" in line) and n < 10:
        out.append(line.rstrip())
        n += 1
print("\n".join(out))
