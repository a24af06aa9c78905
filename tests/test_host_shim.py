"""The C++ host side (supereight_b200/host): DenseSLAMSystem shim + frame-loop driver over the C ABI.
CPU part: it builds and fails loudly without a device.  GPU part: a .raw stream driven through the shim with
the reference's stage gates (benchmark.cpp:115-160) equals the oracle driven through the same gates."""
import os
import struct
import subprocess

import numpy as np
import pytest

import oracle_lib
from supereight_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "supereight_b200", "host")


def exe(field):
    subprocess.run(["make", "-s", "-C", HOST], check=True)
    return os.path.join(HOST, "_build", f"se-denseslam-{field}-b200-benchmark")


def write_stream(tmp, frames, dim, W, H, k, scene="plane"):
    raw = os.path.join(tmp, "scene.raw")
    poses_path = os.path.join(tmp, "poses.txt")
    depth, poses = [], []
    with open(raw, "wb") as f, open(poses_path, "w") as g:
        for i in range(frames):
            d, p = (synth.planar_sweep if scene == "plane" else synth.box_room)(i, dim, W, H, k)
            f.write(struct.pack("<II", W, H)); f.write(d.tobytes())
            f.write(struct.pack("<II", W, H)); f.write(np.zeros((H, W, 3), np.uint8).tobytes())
            g.write(" ".join(repr(float(v)) for v in p.reshape(-1)) + "\n")
            depth.append(d); poses.append(p)
    return raw, poses_path, depth, poses


def read_dump(path, vdtype, loaded=False):
    buf = open(path, "rb").read()
    off, out = 0, []
    extra = (np.uint64, vdtype, np.uint64, vdtype, np.float32, np.float32) if loaded else ()
    for dt in (np.uint64, vdtype, np.uint64, vdtype, np.float32, np.float32, np.uint8, np.uint8, np.uint8) + extra:
        n = struct.unpack_from("<Q", buf, off)[0]; off += 8
        a = np.frombuffer(buf, dtype=dt, count=n, offset=off); off += n * np.dtype(dt).itemsize
        out.append(a)
    return out


def test_host_side_builds_and_fails_loudly_without_gpu(tmp_path):
    from supereight_b200 import capi
    e = exe("sdf")
    assert os.path.exists(e) and os.path.exists(exe("ofusion"))
    if capi.load_library().se_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    raw, poses, _, _ = write_stream(str(tmp_path), 1, 4.8, 160, 120, (120.3, 120.0, 80.0, 60.0))
    r = subprocess.run([e, "-i", raw, "-g", poses, "-k", "120.3,120,80,60"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("field,mu", [("sdf", 0.1), ("ofusion", 0.008)])
def test_shim_frame_loop_matches_oracle(tmp_path, field, mu):
    W, H, size, dim, frames = 160, 120, 256, 4.8, 6
    k = (120.3, 120.0, 80.0, 60.0)
    raw, poses_path, depth, poses = write_stream(str(tmp_path), frames, dim, W, H, k)
    dump = os.path.join(str(tmp_path), "dump.bin")
    log = os.path.join(str(tmp_path), "log.tsv")
    mapfile = os.path.join(str(tmp_path), "test.bin")
    meshfile = os.path.join(str(tmp_path), "mesh.vtk")
    r = subprocess.run([exe(field), "-i", raw, "-g", poses_path, "-v", str(size), "-s", str(dim), "-m", str(mu), "-r", "2", "-z", "1",
                        "-k", ",".join(str(v) for v in k), "-o", log, "-d", dump, "-b", mapfile, "-M", meshfile], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "map file round trip: identical" in r.stderr          # Octree::save format written, loaded into a fresh map, re-exported
    hdr = struct.unpack_from("<ifQ", open(mapfile, "rb").read(16))
    assert hdr[0] == size and abs(hdr[1] - dim) < 1e-6 and hdr[2] > 1
    rows = [ln.split("\t") for ln in open(log).read().strip().split("\n")]
    assert rows[0][0] == "frame" and len(rows) == frames + 1 and len(rows[1]) == 14
    integrated = [int(rw[13]) for rw in rows[1:]]
    assert integrated == [1, 1, 1, 1, 1, 0]                      # rate 2: frames 0..3 always, then even frames (DenseSLAMSystem.cpp:209)

    fid = oracle_lib.SDF if field == "sdf" else oracle_lib.OFUSION
    o = oracle_lib.Oracle(fid, size, dim, W, H)
    for f in range(frames):
        o.preprocess(depth[f])
        if f % 2 == 0 or f <= 3:
            o.integrate(poses[f], k, mu, f)
        if f > 2:
            o.raycast(poses[f], k, mu)
    keys, _, _, data = o.blocks_sorted()
    codes, _, _, values = o.nodes_sorted()
    gk, gd, gc, gv, vert, norm, vol, dep, trk = read_dump(dump, o.vdtype)
    assert np.array_equal(gk, keys) and np.array_equal(gc, codes)
    gd = gd.reshape(-1, 512); gv = gv.reshape(-1, 8)
    vert = vert.reshape(H, W, 3); norm = norm.reshape(H, W, 3)
    # dump_mesh: the file writeVtkMesh writes (commons.h:325-391)
    vtk = open(meshfile).read().split("\n")
    assert vtk[:4] == ["# vtk DataFile Version 1.0", "vtk mesh generated from KFusion", "ASCII", "DATASET POLYDATA"]
    npts = int(vtk[4].split()[1])
    assert vtk[4] == f"POINTS {npts} FLOAT" and npts % 3 == 0 and vtk[5 + npts] == f"POLYGONS {npts // 3} {npts // 3 * 4}"
    assert vtk[6 + npts] == "3 0 1 2" and vtk[5 + npts + npts // 3] == f"3 {npts - 3} {npts - 2} {npts - 1}"
    pts = np.array([ln.split() for ln in vtk[5:5 + npts]], np.float64).reshape(-1, 3, 3)
    if field == "sdf":
        import mc_table_ref
        want = o.marching_cube(mc_table_ref.classic_table())
        assert pts.shape == want.shape and len(want) > 1000
        np.testing.assert_allclose(pts, want, rtol=2e-5, atol=0)      # operator<< prints 6 significant digits
        assert np.array_equal(gd["x"].view(np.uint32), data["x"].view(np.uint32)) and np.array_equal(gd["y"], data["y"])
        assert np.array_equal(gv["x"].view(np.uint32), values["x"].view(np.uint32))
        assert np.array_equal(vert.view(np.uint32), o.vertex().view(np.uint32))
        assert np.array_equal(norm.view(np.uint32), o.normal().view(np.uint32))
        assert np.array_equal(vol.reshape(H, W, 4), o.render_volume(poses[-1], k, mu, 0.75 * mu, False))
    else:
        np.testing.assert_allclose(gd["x"], data["x"], rtol=2e-6, atol=1e-7)      # (1 ulp where glibc log2f is not correctly rounded: tests/test_gpu_parity.py OFU_*)
        assert np.array_equal(gd["y"], data["y"])
    assert np.array_equal(dep.reshape(H, W, 4), o.render_depth())
    assert np.all(trk.reshape(H, W, 4)[..., :3] == np.array([255, 128, 128], np.uint8))     # result 0 -> default colour (rendering.cpp:203-208)


@pytest.mark.gpu
@pytest.mark.parametrize("field,mu", [("sdf", 0.1), ("ofusion", 0.03)])
def test_map_file_exchange_with_the_reference(tmp_path, field, mu):
    """N3 against the reference's own Octree::save / Octree::load (oracle/_ref, octree.hpp:897-950), both directions:
    (1) the map the CUDA kernels built, written by the shim (`-b`), is read by the reference's loader: same octree (as far
        as that loader restores one: see tests/test_map_file.py) -- and the file holds the oracle's payloads byte for byte;
    (2) a map the REFERENCE built and saved is loaded by the shim (`-L`: MapSnapshot::load + setMap) into the device pools:
        re-exported identical, and its raycast is the reference's raycast of that map, bit for bit."""
    if not oracle_lib.have_reference_build():
        pytest.skip("oracle/_ref (the reference build) is absent")
    import test_map_file as tmf
    W, H, size, dim, frames = 160, 120, 256, 4.8, 4
    k = (120.3, 120.0, 80.0, 60.0)
    scene = "plane" if field == "sdf" else "room"
    raw, poses_path, depth, poses = write_stream(str(tmp_path), frames, dim, W, H, k, scene=scene)
    fid = oracle_lib.SDF if field == "sdf" else oracle_lib.OFUSION
    ref = oracle_lib.Oracle(fid, size, dim, W, H, kind=f"ref_{field}")
    ref.lib.seo_set_omp_threads(1)
    for f in range(frames):
        ref.preprocess(depth[f]); ref.integrate(poses[f], k, mu, f)
    ref.raycast(poses[-1], k, mu)
    refmap, shimmap, dump = (os.path.join(str(tmp_path), n) for n in ("reference.bin", "shim.bin", "dump.bin"))
    ref.save_map(refmap)
    r = subprocess.run([exe(field), "-i", raw, "-g", poses_path, "-v", str(size), "-s", str(dim), "-m", str(mu), "-r", "1", "-z", "1",
                        "-k", ",".join(str(v) for v in k), "-o", os.path.join(str(tmp_path), "log.tsv"), "-d", dump, "-b", shimmap, "-L", refmap],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # (1) device map -> shim file -> the reference's reader, and the file's payloads against the reference's own map
    back = oracle_lib.Oracle.load_map(fid, shimmap, f"ref_{field}", dim_fix=dim)
    size_f, dim_f, nodes, blocks = tmf.parse_map_file(shimmap, ref.vdtype)
    keys, coords, _, data = ref.blocks_sorted()
    codes, side, _, values = ref.nodes_sorted()
    assert (size_f, np.float32(dim_f)) == (size, np.float32(dim)) and len(keys) > 300
    assert np.array_equal(blocks["code"], keys) and np.array_equal(blocks["coords"], coords)
    assert np.array_equal(nodes["code"], codes) and np.array_equal(nodes["side"], side)
    if field == "sdf":
        assert blocks["voxels"].tobytes() == data.tobytes() and nodes["value"].tobytes() == values.tobytes()
        tmf.same_octree(ref, back, reloaded_by_reference=True)
    else:
        np.testing.assert_allclose(blocks["voxels"]["x"], data["x"], rtol=2e-6, atol=1e-7)
        assert np.array_equal(blocks["voxels"]["y"], data["y"])
        kb, cb, _, _ = back.blocks_sorted(with_data=False)
        assert np.array_equal(kb, keys) and np.array_equal(cb, coords) and np.array_equal(back.nodes_sorted()[0], codes)
    back.close()
    # (2) reference map -> file -> shim -> device: content and raycast
    out = read_dump(dump, ref.vdtype, loaded=True)
    lk, ld, lc, lv, lvert, lnorm = out[9:]
    assert np.array_equal(lk, keys) and np.array_equal(lc, codes)
    assert ld.tobytes() == data.tobytes() and lv.tobytes() == values.tobytes()
    lvert = lvert.reshape(H, W, 3); lnorm = lnorm.reshape(H, W, 3)
    rv, rn = ref.vertex(), ref.normal()
    assert (rn[..., 0] != -2.0).sum() > 0.2 * W * H
    if field == "sdf":
        assert np.array_equal(lvert.view(np.uint32), rv.view(np.uint32)) and np.array_equal(lnorm.view(np.uint32), rn.view(np.uint32))
    else:
        hit_g, hit_r = lnorm[..., 0] != -2.0, rn[..., 0] != -2.0
        assert (hit_g != hit_r).mean() <= 2e-5
        both = hit_g & hit_r
        np.testing.assert_allclose(lvert[both], rv[both], rtol=0, atol=2e-6)
