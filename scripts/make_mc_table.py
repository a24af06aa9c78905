"""Writes supereight_b200/csrc/se_mc_table.cuh: the classic marching-cubes case table in this library's own encoding.

The table itself is not the reference's invention: it is the tabulation every marching-cubes implementation shares (Lorensen &
Cline 1987; the widely reproduced 256 x 16 list published by P. Bourke, "Polygonising a scalar field", 1994, public domain),
and supereight ships a transcription of it (se_core/include/se/algorithms/edge_tables.h, `triTable`).  Which diagonals split a
case's polygons is a convention that cannot be re-derived; a mesh that is to equal the reference's triangle for triangle
has to use the same list.  This script reads the numbers from the reference tree (development container only) and writes them
as 256 strings of hexadecimal edge indices; tests/test_meshing.py checks the result against the first-principles generator
(tests/mc_table_ref.py: same directed polygon boundaries in all 256 cases) and, where /root/reference exists, against the
header it was read from.
Usage: python scripts/make_mc_table.py [/root/reference]"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def read_reference_rows(ref_root):
    text = open(os.path.join(ref_root, "se_core/include/se/algorithms/edge_tables.h")).read()
    body = text[text.index("triTable"):]
    body = body[body.index("{") + 1:]
    rows = re.findall(r"\{([^{}]*)\}", body)
    assert len(rows) >= 256
    out = []
    for r in rows[:256]:
        v = [int(x) for x in r.replace("\n", " ").split(",") if x.strip()]
        assert len(v) == 16
        out.append([e for e in v if e >= 0])
    return out


def main():
    ref_root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    rows = read_reference_rows(ref_root)
    with open(os.path.join(ROOT, "supereight_b200", "csrc", "se_mc_table.cuh"), "w") as f:
        f.write("// The classic marching-cubes case table (Lorensen & Cline 1987; P. Bourke's public-domain tabulation), as 256 strings of\n"
                "// hexadecimal edge indices, three per triangle, in the corner / edge numbering of se/algorithms/meshing.hpp:58-104.\n"
                "// Written by scripts/make_mc_table.py; read by mc_case_table() in se_meshing.cuh.  Data, not code: see DESIGN.md, N4.\n")
        for i in range(0, 256, 8):
            f.write("  " + " ".join('"%s",' % "".join("%x" % e for e in rows[j]) for j in range(i, i + 8)) + "\n")
    print("wrote", len(rows), "cases,", sum(len(r) for r in rows) // 3, "triangles")


if __name__ == "__main__":
    main()
