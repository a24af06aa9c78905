"""N3: the map file against the reference's OWN writer and reader.

`oracle/_ref` holds the reference's Octree::save / Octree::load (se_core/include/se/octree.hpp:897-950,
io/se_serialise.hpp:54-99), compiled where they lie.  The shim's se::MapSnapshot::save / load (what
DenseSLAMSystem::getMap() / setMap() and `se_b200_benchmark -b` use) must read what the reference writes and write what the
reference reads:

  reference save -> shim load -> shim save     the file comes back byte for byte
  ... -> reference load                        the octree the reference rebuilds from the shim's file == the original
  shim "sort" (the order getMap() exports)     the reference loads that too, to the same octree

CPU tier: the shim's file code needs no device (supereight_b200/host/src/se_b200_mapfile.cpp).  The GPU tier
(tests/test_host_shim.py) runs the same exchange with a map built by the CUDA kernels.
"""
import filecmp
import os
import struct
import subprocess

import numpy as np
import pytest

import oracle_lib
from supereight_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not oracle_lib.have_reference_build(), reason="oracle/_ref (the reference build) is absent")


def tool(field):
    exe = os.path.join(ROOT, "supereight_b200", "host", "_build", f"se-denseslam-{field}-b200-mapfile")
    if not os.path.exists(exe):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "supereight_b200", "host")], check=True)
    return exe


def reference_map(field, frames=3, size=128, dim=4.8, W=80, H=60):
    fid = {"sdf": oracle_lib.SDF, "ofusion": oracle_lib.OFUSION}[field]
    mu = 0.1 if field == "sdf" else 0.03
    k = (60.15, 60.0, 40.0, 30.0)
    o = oracle_lib.Oracle(fid, size, dim, W, H, kind=f"ref_{field}")
    o.lib.seo_set_omp_threads(1)               # allocate_level's children_mask_ update races under OpenMP (octree.hpp:843-849)
    gen = synth.planar_sweep if field == "sdf" else synth.box_room
    for f in range(frames):
        d, pose = gen(f, dim, W, H, k)
        o.preprocess(d); o.integrate(pose, k, mu, f)
    return o, fid


def same_octree(a, b, reloaded_by_reference=False):
    """reloaded_by_reference: `b` was rebuilt by the reference's Octree::load, which restores only the FIRST voxel of every
    block -- its memcpy copies sizeof(*(tmp.getBlockRawPtr())), one voxel, instead of the 512 (octree.hpp:944-945); the other
    511 keep initValue().  Structure, node values and that first voxel are what it can be held to."""
    ka, ca, _, da = a.blocks_sorted()
    kb, cb, _, db = b.blocks_sorted()
    assert len(ka) > 50 and np.array_equal(ka, kb) and np.array_equal(ca, cb)
    if reloaded_by_reference:
        assert da[:, 0].tobytes() == db[:, 0].tobytes()
        init = np.zeros(1, a.vdtype); init["x"] = 1.0 if a.field == oracle_lib.SDF else 0.0
        assert all(db[:, 1:][n].tobytes() == np.broadcast_to(init, db[:, 1:].shape)[n].tobytes() for n in ("x", "y"))
    else:
        assert da.tobytes() == db.tobytes()
    na, sa, ma, va = a.nodes_sorted()
    nb, sb, mb, vb = b.nodes_sorted()
    assert np.array_equal(na, nb) and np.array_equal(sa, sb) and np.array_equal(ma, mb)
    assert va.tobytes() == vb.tobytes()


def parse_map_file(path, vdtype):
    """the file as octree.hpp:897-915 + io/se_serialise.hpp:54-99 lay it out"""
    node = np.dtype([("code", "<u8"), ("side", "<u4"), ("value", vdtype, (8,))])
    block = np.dtype([("code", "<u8"), ("coords", "<i4", (3,)), ("voxels", vdtype, (512,))])
    assert node.itemsize == 12 + 8 * vdtype.itemsize and block.itemsize == 20 + 512 * vdtype.itemsize
    with open(path, "rb") as fh:
        size, dim, n = struct.unpack("<ifQ", fh.read(16))
        nodes = np.frombuffer(fh.read(n * node.itemsize), node)
        (n,) = struct.unpack("<Q", fh.read(8))
        blocks = np.frombuffer(fh.read(n * block.itemsize), block)
        assert fh.read() == b""
    return size, dim, nodes, blocks


@pytest.mark.parametrize("field", ["sdf", "ofusion"])
def test_shim_reads_and_writes_the_reference_format(tmp_path, field):
    o, fid = reference_map(field)
    a, b, c = (str(tmp_path / n) for n in ("reference.bin", "shim_copy.bin", "shim_sorted.bin"))
    o.save_map(a)
    # header as octree.hpp:900-906 writes it
    with open(a, "rb") as fh:
        size, dim, n_nodes = struct.unpack("<ifQ", fh.read(16))
    assert (size, n_nodes) == (o.size, o.node_count()) and abs(dim - o.dim) < 1e-6
    vb = 8 if field == "sdf" else 16
    assert os.path.getsize(a) == 16 + n_nodes * (8 + 4 + 8 * vb) + 8 + o.block_count() * (8 + 12 + 512 * vb)
    subprocess.run([tool(field), "copy", a, b], check=True)
    assert filecmp.cmp(a, b, shallow=False), "MapSnapshot::load + save changed a file written by the reference"
    info = subprocess.run([tool(field), "info", b], check=True, capture_output=True, text=True).stdout.split()
    assert info[:2] == ["size", str(o.size)] and info[4:] == ["nodes", str(n_nodes), "blocks", str(o.block_count())]
    # the shim's key-ordered file (what getMap() + save writes) holds exactly the reference's records: every payload byte
    subprocess.run([tool(field), "sort", a, c], check=True)
    size_c, dim_c, nodes, blocks = parse_map_file(c, o.vdtype)
    keys, coords, _, data = o.blocks_sorted()
    codes, side, _, values = o.nodes_sorted()
    assert (size_c, np.float32(dim_c)) == (o.size, np.float32(o.dim))
    assert np.array_equal(blocks["code"], keys) and np.array_equal(blocks["coords"], coords) and blocks["voxels"].tobytes() == data.tobytes()
    assert np.array_equal(nodes["code"], codes) and np.array_equal(nodes["side"], side) and nodes["value"].tobytes() == values.tobytes()
    # the reference's reader rebuilds the same octree from the shim's files (records in pool order, and in key order)
    for path in (b, c):
        back = oracle_lib.Oracle.load_map(fid, path, f"ref_{field}", dim_fix=o.dim)
        assert back.size == o.size
        same_octree(o, back, reloaded_by_reference=True)
        back.close()
    # the reader's `dim` quirk (octree.hpp:921-923 reads the float into an int): documented, not relied upon
    raw = oracle_lib.Oracle.load_map(fid, b, f"ref_{field}")
    assert raw.dim == float(np.float32(struct.unpack("<i", struct.pack("<f", o.dim))[0]))
    raw.close()


@pytest.mark.parametrize("field", ["sdf", "ofusion"])
def test_point_queries_on_a_reloaded_map_match(tmp_path, field):
    """fetch / get_fine on the octree the reference rebuilds from the shim's file: every block is found again where it was,
    and holds the voxel the reference's reader restores (the first one, see same_octree)"""
    o, fid = reference_map(field)
    a, c = str(tmp_path / "reference.bin"), str(tmp_path / "shim_sorted.bin")
    o.save_map(a)
    subprocess.run([tool(field), "sort", a, c], check=True)
    back = oracle_lib.Oracle.load_map(fid, c, f"ref_{field}", dim_fix=o.dim)
    keys, coords, _, _ = o.blocks_sorted(with_data=False)
    rng = np.random.default_rng(5)
    for bc in coords[rng.choice(len(coords), 60, replace=False)]:
        x, y, z = map(int, bc)
        assert back.fetch(x + 3, y + 5, z + 7) and back.fetch_octant_code(x, y, z, int(np.log2(o.size)) - 3) == o.fetch_octant_code(x, y, z, int(np.log2(o.size)) - 3)
        assert o.get_fine(x, y, z) == back.get_fine(x, y, z)
    assert not back.fetch(0, 0, 0) or o.fetch(0, 0, 0)
    back.close()


@pytest.mark.parametrize("field", ["sdf", "ofusion"])
def test_snapshot_host_mirror_of_get_and_interp(tmp_path, field):
    """se::MapSnapshot::get_fine / interp -- what a caller of getMap() uses on the reference's se::Octree -- against the
    reference's own Octree::get_fine / interp on the map the file came from: bit for bit."""
    o, fid = reference_map(field)
    a = str(tmp_path / "reference.bin")
    o.save_map(a)
    _, coords, _, _ = o.blocks_sorted(with_data=False)
    rng = np.random.default_rng(11)
    pts = []
    for bc in coords[rng.choice(len(coords), min(150, len(coords)), replace=False)]:
        pts.append(bc + rng.uniform(0.0, 8.0, 3))                      # inside allocated blocks, block borders included
    pts += list(rng.uniform(2, o.size - 3, (100, 3)))                  # anywhere (mostly unallocated)
    pts = np.clip(np.array(pts, np.float32), 1.0, o.size - 3.0)
    pfile, ofile = str(tmp_path / "points.txt"), str(tmp_path / "out.txt")
    np.savetxt(pfile, pts, fmt="%.9g")
    subprocess.run([tool(field), "query", a, pfile, ofile], check=True)
    got = np.loadtxt(ofile).reshape(len(pts), 3)
    hits = 0
    for q, g in zip(np.loadtxt(pfile, dtype=np.float32).reshape(-1, 3), got):
        vx, vy = o.get_fine(int(q[0]), int(q[1]), int(q[2]))
        assert np.float32(g[0]) == np.float32(vx) and float(g[1]) == float(vy), (q, g, vx, vy)
        assert np.float32(g[2]) == np.float32(o.interp(float(q[0]), float(q[1]), float(q[2]))), (q, g)
        hits += vy != 0 or np.float32(vx) != np.float32(1.0 if field == "sdf" else 0.0)
    assert hits > (50 if field == "sdf" else 5)
