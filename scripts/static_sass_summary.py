"""Static summary of the built library's device code (build container, no GPU): per kernel the registers / stack / shared
memory (`cuobjdump -res-usage`), the instruction count and the sm_100a-specific opcodes (`cuobjdump -sass`).
Usage: python scripts/static_sass_summary.py [lib.so] > profiles/<round>_static_sass.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "supereight_b200", "libse_b200.so")
KEY = ["UBLKCP", "SYNCS", "ACQBULK", "PREEXIT", "FFMA2", "FMUL2", "FADD2", "MATCH", "VOTE", "ATOMG", "RED", "LDS", "STS", "LDG", "STG", "MUFU", "FCHK", "CALL"]


def demangle(names):
    return subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.splitlines()


res_text = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True, check=True).stdout.splitlines()
res = {}
for i, ln in enumerate(res_text):
    m = re.match(r"\s*Function (\S+):", ln)
    if m:
        res[m.group(1)] = {k: int(v) for k, v in re.findall(r"(\w+)(?:\[0\])?:(\d+)", res_text[i + 1])}
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
per, cur = collections.OrderedDict(), None
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = per.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur is not None:
        cur[m.group(1)] += 1
names = list(per)
pretty = dict(zip(names, demangle(names)))
print(f"# static device-code summary of {os.path.relpath(lib, ROOT)} (cuobjdump -res-usage / -sass; sm_100a)")
print("# kernel | REG STACK SHARED(static) | instructions | " + " ".join(KEY))
for n in sorted(names, key=lambda k: pretty[k]):
    ops = per[n]
    short = re.sub(r"^void ", "", pretty[n]).split("(")[0].replace("se_b200::", "")
    r = res.get(n, {})
    keys = " ".join(f"{k}:{sum(v for o, v in ops.items() if o.startswith(k))}" for k in KEY if any(o.startswith(k) for o in ops))
    print(f"{short} | {r.get('REG', '?')} {r.get('STACK', '?')} {r.get('SHARED', '?')} | {sum(ops.values())} | {keys}")
