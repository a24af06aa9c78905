"""Static SASS instruction count per source line of one kernel (build container, no GPU needed).
Usage: python scripts/sass_by_line.py <kernel-name-substring> [lib.so] [min_count]
Extracts the cubin from the library, disassembles it with line info (`nvdisasm -g`; the library is built with -lineinfo)
and prints, for the first kernel whose mangled name contains the substring: total instructions, and file:line -> count."""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def by_line(kernel_substr, lib):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, check=True, capture_output=True)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        text = subprocess.run(["nvdisasm", "-g", cubin], cwd=d, check=True, capture_output=True, text=True).stdout
    counts, cur, inside, name = collections.Counter(), None, False, None
    for line in text.splitlines():
        if line.startswith(".text."):
            if inside:
                break
            inside = kernel_substr in line
            name = line[6:-1] if inside else name
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        elif re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
            counts[cur] += 1
    return name, counts


if __name__ == "__main__":
    lib = os.path.abspath(sys.argv[2]) if len(sys.argv) > 2 else os.path.join(ROOT, "supereight_b200", "libse_b200.so")
    floor = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    name, counts = by_line(sys.argv[1], lib)
    print(name, "total", sum(counts.values()))
    for (f, l), n in sorted(counts.items(), key=lambda kv: (kv[0] is None, kv[0])):
        if n >= floor:
            print(f"{f}:{l}\t{n}")
