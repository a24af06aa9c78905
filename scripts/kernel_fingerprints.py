"""Per-kernel SHA-256 of the library's SASS (build container, no GPU).
  python scripts/kernel_fingerprints.py [lib.so] > profiles/<name>.sha256        write
  python scripts/kernel_fingerprints.py --check profiles/<name>.sha256 [lib.so]  compare: which kernels changed / are new / are gone
profiles/r1b_gpu_verified_kernels.sha256 is the build that passed the 31 GPU parity tests and was benchmarked in the round's last
successful GPU call (profiles/r1b_ab_grad_rows_ofusion_fast.log): a later library whose kernels all match it runs the same device code."""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fingerprints(lib):
    text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    per, cur = collections.OrderedDict(), None
    for ln in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = per.setdefault(m.group(1), hashlib.sha256())
            continue
        if cur is not None and re.match(r"\s*/\*[0-9a-f]{4}\*/", ln):
            cur.update(ln.strip().encode())
    return {k: v.hexdigest() for k, v in per.items()}


if __name__ == "__main__":
    args = sys.argv[1:]
    if args and args[0] == "--check":
        want = dict(line.split()[::-1] for line in open(args[1]) if line.strip() and not line.startswith("#"))
        got = fingerprints(args[2] if len(args) > 2 else os.path.join(ROOT, "supereight_b200", "libse_b200.so"))
        changed = sorted(k for k in want if k in got and got[k] != want[k])
        print(f"{len(want)} recorded, {len(got)} in the library; changed: {changed or 'none'}; new: {sorted(set(got) - set(want)) or 'none'}; "
              f"gone: {sorted(set(want) - set(got)) or 'none'}")
        sys.exit(1 if changed or set(want) - set(got) else 0)
    lib = args[0] if args else os.path.join(ROOT, "supereight_b200", "libse_b200.so")
    for k, v in sorted(fingerprints(lib).items()):
        print(v, k)
