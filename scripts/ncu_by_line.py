"""Dynamic per-source-line profile of one kernel from an .ncu-rep captured with --import-source on (build container).
Usage: python scripts/ncu_by_line.py <report.ncu-rep> <kernel-regex> [min_pct]
Prints, per source file:line, executed warp-instructions, their share, stall samples and average active threads."""
import collections
import csv
import io
import os
import subprocess
import sys


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    floor = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", f"regex:{kernel}"],
                         capture_output=True, text=True).stdout
    # the page is a sequence of per-file blocks: "File Path", "Function Name", header row, rows; source rows have a line number,
    # SASS rows under them have an address
    per_line = collections.OrderedDict()
    cur_file, header, cur_line = None, None, None
    total = 0
    for row in csv.reader(io.StringIO(out)):
        if not row:
            continue
        if row[0] == "File Path":
            cur_file = os.path.basename(row[1]); header = None; continue
        if row[0] == "Function Name":
            continue
        if row[0] == "Line No":
            header = row; continue
        if header is None:
            continue
        d = dict(zip(header, row))
        if row[0] != "":
            cur_line = (cur_file, int(row[0]), row[1].strip()[:90])
            continue
        def num(key):
            try:
                return int(d.get(key) or 0)
            except ValueError:
                return 0
        ex, th, smp = num("Instructions Executed"), num("Thread Instructions Executed"), num("# Samples")
        e = per_line.setdefault(cur_line, [0, 0, 0])
        e[0] += ex; e[1] += th; e[2] += smp
        total += ex
    tot_smp = sum(v[2] for v in per_line.values()) or 1
    print(f"kernel {kernel}: {total} warp-instructions, {tot_smp} samples")
    byfile = collections.Counter()
    for (f, l, src), (ex, th, smp) in per_line.items():
        byfile[f] += ex
        if 100.0 * ex / total >= floor or 100.0 * smp / tot_smp >= floor:
            print(f"{f}:{l:<5d} {100.0 * ex / total:5.1f}% instr {100.0 * smp / tot_smp:5.1f}% smp  thr {th / max(ex, 1):4.1f}  | {src}")
    print({k: round(100.0 * v / total, 1) for k, v in byfile.items()})


if __name__ == "__main__":
    main()
