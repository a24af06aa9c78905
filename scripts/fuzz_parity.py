"""Randomised parity check of the product kernels against the oracle: random volumes, cameras (inside and outside the volume,
rolled, negative fy), depth images (planes, blobs, noise, drop-outs, saturated and 1 mm samples) and short sequences; after
every scenario the block set, voxel and node values, vertex / normal maps and both renderings are compared -- bit for bit
for the SDF field, to the test suite's tolerances for OFusion.
  python scripts/fuzz_parity.py [n_scenarios] [first_seed]
  python scripts/fuzz_parity.py --ref <size> [n_scenarios] [first_seed]
  python scripts/fuzz_parity.py --ref-device <size> [n_scenarios] [first_seed]
With --ref-device the PRODUCT is compared directly with the reference's own code over the --ref scenario domain (the reference's
defined domain, see below): one hop instead of product == oracle == reference.
With --ref the same scenarios compare the ORACLE with the reference's own code (oracle/_ref, built where /root/reference exists)
instead of the product with the oracle -- every array bit for bit, both fields, overflowing allocation lists included (single
thread: the truncation is then deterministic).  One volume size per process: the reference keeps a `static const float epsilon`
of the first map it sees (ray_iterator.hpp:63).
Runs wherever the library runs: on a B200, or in the build container on the fiber executor
(SE_B200_LIB=tests/simt_emu/_build/libse_b200_simt.so; ~1 s per scenario).  Prints one line per failing scenario and a summary;
exit status 1 if anything differed."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib  # noqa: E402
from oracle_lib import OFUSION, SDF, Oracle  # noqa: E402
from parity_utils import compare_blocks, compare_images, compare_nodes  # noqa: E402

import mc_table_ref  # noqa: E402  (tests/: reader of the marching-cubes case table the library ships)
from supereight_b200 import Map, synth  # noqa: E402

MC_TABLE = mc_table_ref.classic_table()


class RefAsMap:
    """the reference build behind the few Map methods a scenario uses"""
    def __init__(self, field, size, dim, W, H):
        self.o = Oracle(field, size, dim, W, H, kind="ref_sdf" if field == SDF else "ref_ofusion")
        self.o.lib.seo_set_omp_threads(1)      # allocate_level's children_mask_ update is racy otherwise (octree.hpp:843-849)
    def __getattr__(self, name):
        return getattr(self.o, name)
    def vertex_normal(self):
        return self.o.vertex(), self.o.normal()
    def query_interp(self, pos):
        return np.array([self.o.interp(float(q[0]), float(q[1]), float(q[2])) for q in pos], np.float32)
    def query_grad(self, pos):
        return np.array([self.o.grad(float(q[0]), float(q[1]), float(q[2])) for q in pos], np.float32)


def rays_reach_a_face(o, view, k, dim, pixels, far):
    """True when every listed pixel's ray leaves the volume before `far` metres: the march then samples the last voxel slice, where
    the reference's interp / grad read beyond the volume (undefined; the oracle and the library read initValue() there)."""
    fx, fy, cx, cy = [np.float32(v) for v in k]
    Kinv = np.eye(4, dtype=np.float32)
    Kinv[0, 0] = np.float32(1) / fx; Kinv[1, 1] = np.float32(1) / fy; Kinv[0, 2] = -cx / fx; Kinv[1, 2] = -cy / fy
    V = (np.asarray(view, np.float32) @ Kinv).astype(np.float32)
    org = np.asarray(view, np.float32)[:3, 3]
    for x, y in pixels:
        d = V[:3, :3] @ np.array([x, y, 1], np.float32)
        d = (d / np.sqrt(np.float32((d * d).sum()))).astype(np.float32)
        _, tinfo = o.ray_blocks(org, d, 0.4, far)
        if not tinfo[1] < np.float32(far) * np.float32(0.999):
            return False
    return True


def inside_pose(rng, dim):
    """camera well inside the volume, looking at a point inside (--ref mode: the reference's defined domain)"""
    eye = rng.uniform(0.25, 0.75, 3) * dim
    tgt = rng.uniform(0.2, 0.8, 3) * dim
    if np.linalg.norm(tgt - eye) < 0.05 * dim:
        tgt = eye + np.array([0.0, 0.0, 0.2 * dim])
    pose = synth.look_at_pose(eye, tgt, rng.uniform(-np.pi, np.pi))
    return pose if np.all(np.isfinite(pose)) else synth.yaw_pose(*eye, 0.3)


def keep_surfaces_inside(d_mm, pose, k, dim, margin):
    """zeroes / pulls in every depth sample whose surface point would lie within `margin` metres of a volume face: beyond the
    faces the reference indexes out of bounds (interp / grad of the last voxel slice, rays that start outside the volume)"""
    H, W = d_mm.shape
    dw, t = synth._rays(W, H, k, pose)                         # surface point = t + depth * dw   (depth along the optical axis)
    with np.errstate(divide="ignore", invalid="ignore"):
        lim = np.where(dw > 0, (dim - margin - t) / dw, np.where(dw < 0, (margin - t) / dw, np.inf))
    dmax = np.clip(lim.min(axis=-1), 0.0, 65.0)                # metres of depth before the ray leaves the shrunken volume
    out = np.minimum(d_mm.astype(np.float64), np.floor(dmax * 1000.0))
    out[dmax < 0.05] = 0
    return np.ascontiguousarray(out.astype(np.uint16))


def random_pose(rng, dim):
    mode = rng.integers(4)
    if mode == 0:        # inside, looking at a random point inside
        eye = rng.uniform(0.05, 0.95, 3) * dim
    elif mode == 1:      # outside the volume, looking in
        eye = rng.uniform(-0.6, 1.6, 3) * dim
    elif mode == 2:      # on a face / edge of the volume
        eye = rng.uniform(0.0, 1.0, 3) * dim
        eye[rng.integers(3)] = rng.choice([0.0, dim])
    else:                # axis-aligned view (zero direction components exercise the epsilon clamp of the ray set-up)
        eye = rng.uniform(0.1, 0.9, 3) * dim
        tgt = eye.copy(); tgt[rng.integers(3)] += rng.choice([-1.0, 1.0]) * dim
        with np.errstate(invalid="ignore", divide="ignore"):
            pose = synth.look_at_pose(eye, tgt, 0.0)              # looking straight along y has no defined "up": NaN
        if not np.all(np.isfinite(pose)):
            pose = synth.yaw_pose(*eye, float(rng.choice([0.0, np.pi / 2, np.pi])))
        return pose
    tgt = rng.uniform(0.2, 0.8, 3) * dim
    if np.linalg.norm(tgt - eye) < 1e-3:
        tgt = tgt + 0.1 * dim
    pose = synth.look_at_pose(eye, tgt, rng.uniform(-np.pi, np.pi))
    if not np.all(np.isfinite(pose)):
        pose = synth.yaw_pose(*(rng.uniform(0.2, 0.8, 3) * dim), rng.uniform(-3, 3))
    return pose


def random_depth(rng, W, H, dim):
    kind = rng.integers(4)
    u, v = np.meshgrid(np.arange(W), np.arange(H))
    if kind == 0:        # tilted plane
        d = rng.uniform(0.3, 1.2) * dim * (1 + rng.uniform(-0.5, 0.5) * (u / W - 0.5) + rng.uniform(-0.5, 0.5) * (v / H - 0.5))
    elif kind == 1:      # blobs on a background
        d = np.full((H, W), rng.uniform(0.5, 1.5) * dim)
        for _ in range(rng.integers(1, 5)):
            cx, cy, r = rng.uniform(0, W), rng.uniform(0, H), rng.uniform(3, W / 3)
            d[(u - cx) ** 2 + (v - cy) ** 2 < r * r] = rng.uniform(0.1, 1.0) * dim
    elif kind == 2:      # white noise between near and far
        d = rng.uniform(0.05, 1.5, (H, W)) * dim
    else:                # steps
        d = (1 + (u // max(W // 6, 1)) % 3) * rng.uniform(0.15, 0.4) * dim
    d = d * 1000.0 + rng.normal(0, rng.choice([0.0, 2.0, 20.0]), (H, W))
    out = np.clip(np.rint(d), 0, 65535).astype(np.uint16)
    out[rng.random((H, W)) < rng.choice([0.0, 0.02, 0.3])] = 0
    if rng.random() < 0.3:
        out[rng.random((H, W)) < 0.02] = 65535        # saturated samples
    if rng.random() < 0.3:
        out[rng.random((H, W)) < 0.02] = 1            # 1 mm samples
    return np.ascontiguousarray(out)


def scenario(seed, ref_size=0, device_vs_ref=False):
    rng = np.random.default_rng(seed)
    field = int(rng.integers(2))
    size = int(rng.choice([16, 32, 64, 128, 256]))
    if ref_size:
        size = ref_size
    dim = float(rng.choice([0.5, 1.0, 2.0, 4.8, 10.0]))
    W, H = int(rng.integers(9, 97)), int(rng.integers(5, 73))
    f = rng.uniform(0.6, 2.0) * W
    k = (float(f), float(f * rng.choice([1.0, -1.0, 0.9])), float(W / 2 + rng.uniform(-5, 5)), float(H / 2 + rng.uniform(-5, 5)))
    mu = float(rng.choice([0.1, 0.05, 0.02]) if field == SDF else rng.choice([0.008, 0.02]))
    if field == SDF and 2 * mu / (dim / size) > 90:
        mu = 40 * dim / size                              # keeps the band below the per-ray block list (100 samples)
    g, o = (RefAsMap if ref_size and not device_vs_ref else Map)(field, size, dim, W, H), Oracle(field, size, dim, W, H)
    ref = RefAsMap(field, size, dim, W, H) if device_vs_ref else None      # --ref-device: product vs reference; the oracle only gates the domain
    o.set_counting(True)
    reserved = (size // 8) * W * H             # DenseSLAMSystem.cpp:212-215
    n_frames = int(rng.integers(1, 5))
    pose = None
    for fr in range(n_frames):
        if pose is None or rng.random() < 0.5:
            pose = inside_pose(rng, dim) if ref_size else random_pose(rng, dim)
        d = random_depth(rng, W, H, dim)
        if ref_size:
            d = keep_surfaces_inside(d, pose, k, dim, (6 if field == OFUSION else 2) * mu + 4 * dim / size)
        o.reset_counters()
        o.preprocess(d); o.integrate(pose, k, mu, fr)
        if ref_size and o.counters()["n_keys_raw"] == 0:
            return None        # allocate(keys, 0): the reference processes one stale key of an earlier frame (unique.hpp:51-60), undefined
        if o.counters()["n_keys_raw"] >= reserved and (not ref_size or device_vs_ref):
            # The reference stops recording requests when its reserved list is full (alloc_impl.hpp:103-106); which requests
            # are lost depends on the OpenMP interleaving, so there is no reference answer.  (The library has no such list.)
            return None
        g.preprocess(d); g.integrate(pose, k, mu, fr)
        if ref is not None:
            ref.preprocess(d); ref.integrate(pose, k, mu, fr)
    problems = []
    if ref is not None:
        o = ref
    exact = field == SDF or (ref_size and not device_vs_ref)               # OFusion on the device: log2f is 1 ulp off glibc's on some arguments
    cb = compare_blocks(g, o)
    if not cb["keys_equal"]:
        return [f"block sets differ {cb}"]
    if not cb["coords_equal"] or cb["active_mismatch"]:
        problems.append(f"block coords/active {cb}")
    cn = compare_nodes(g, o)
    if not (cn["codes_equal"] and cn.get("side_equal") and cn.get("mask_equal")):
        problems.append(f"nodes {cn}")
    if exact:
        if cb["x_bit_mismatch"] or cb["y_mismatch"] or cn.get("x_bit_mismatch") or cn.get("y_mismatch"):
            problems.append(f"SDF values {cb} {cn}")
    else:
        # occupancies: the test suite's bar (tests/test_gpu_parity.py assert_ofusion_parity: rtol 1e-4, atol 1e-5) -- the oracle's
        # log2f (glibc) and the device's correctly rounded log2 differ by an ulp on some arguments, and updates accumulate
        if cb["y_mismatch"] or cb["x_max_abs"] > 1e-5 + 1e-4 * 1000.0 or (cb["x_max_rel"] > 1e-4 and cb["x_max_abs"] > 1e-5):
            problems.append(f"OFusion values {cb}")
        if cn.get("y_mismatch") or (cn.get("x_max_rel", 0) > 1e-4):
            problems.append(f"OFusion node values {cn}")
    view = pose if rng.random() < 0.7 else (inside_pose(rng, dim) if ref_size else random_pose(rng, dim))
    o.raycast(view, k, mu); g.raycast(view, k, mu)
    gv, gn = g.vertex_normal()
    ci = compare_images(gv, gn, o.vertex(), o.normal())
    at_face = np.zeros((H, W), bool)
    if ref_size:
        # hits within two voxels of a volume face: interp / grad of the reference read beyond the volume there (undefined)
        ov_, rv_ = o.vertex() * (size / dim), gv * (size / dim)
        for v_, n_ in ((ov_, o.normal()), (rv_, gn)):
            at_face |= (n_[..., 0] != -2.0) & ((v_ < 2.0) | (v_ > size - 3.0)).any(axis=-1)
    if exact:
        differs = (gv.view(np.uint32) != o.vertex().view(np.uint32)).any(axis=-1) | (gn.view(np.uint32) != o.normal().view(np.uint32)).any(axis=-1)
        if (differs & ~at_face).any():
            problems.append(f"raycast {ci}")
        for rer in (False, True):
            a, b = g.render_volume(view, k, mu, 0.75 * mu, rer), o.render_volume(view, k, mu, 0.75 * mu, rer)
            if not np.array_equal(a, b):
                ys, xs = np.nonzero((a != b).any(axis=2))
                if ref_size and not rer and at_face[ys, xs].all():
                    continue       # shading of hits at a volume face, see above
                if ref_size and rer and len(ys) <= 0.05 * W * H and rays_reach_a_face(o, view, k, dim, list(zip(xs, ys)), 8.0):
                    continue       # the reference's undefined reads at the volume face (arithmetic contract)
                problems.append(f"render_volume(reraycast={rer}) differs in {len(ys)} pixels")
        # point queries at random positions (also on and beyond the volume's faces): get / interp / grad, bit for bit
        pos = np.concatenate([rng.uniform(-2, size + 2, (200, 3)), rng.uniform(0, size, (200, 3))]).astype(np.float32)
        if ref_size:
            pos = pos[np.all((pos >= 1) & (pos < size - 2), axis=1)]      # beyond that the reference's interp / grad index out of bounds
        gi, gg = g.query_interp(pos), g.query_grad(pos).reshape(-1, 3)
        oi = np.array([o.interp(float(q[0]), float(q[1]), float(q[2])) for q in pos], np.float32)
        og = np.array([o.grad(float(q[0]), float(q[1]), float(q[2])) for q in pos], np.float32).reshape(-1, 3)
        inside = np.all((pos >= 0) & (pos < size - 1), axis=1)              # (the oracle defines interp beyond the faces; compared inside only)
        if not np.array_equal(gi.view(np.uint32)[inside], oi.view(np.uint32)[inside]):
            problems.append("interp differs")
        if not np.array_equal(gg.view(np.uint32), og.view(np.uint32)):
            problems.append("grad differs")
        # N4: the mesh, triangle by triangle (the reference meshes with its own case table: not compared in --ref mode)
        if not ref_size:
            got, want = g.mesh(), o.marching_cube(MC_TABLE)
            if got.shape != want.shape or not np.array_equal(got.view(np.uint32), want.view(np.uint32)):
                problems.append(f"mesh differs ({got.shape} vs {want.shape})")
    else:
        if ci["hit_mask_mismatch"] > 0.01 * W * H + 2:
            problems.append(f"OFusion raycast {ci}")
    return problems


if __name__ == "__main__":
    argv = sys.argv[1:]
    ref_size, device_vs_ref = 0, False
    if argv and argv[0] in ("--ref", "--ref-device"):
        device_vs_ref = argv[0] == "--ref-device"
        ref_size, argv = int(argv[1]), argv[2:]
        if not oracle_lib.have_reference_build():
            sys.exit("oracle/_ref is not built (it needs /root/reference: `make -C oracle ref`)")
    n = int(argv[0]) if len(argv) > 0 else 100
    first = int(argv[1]) if len(argv) > 1 else 0
    oracle_lib.build()
    bad = skipped = 0
    for s in range(first, first + n):
        try:
            p = scenario(s, ref_size, device_vs_ref)
        except Exception as e:                       # an error return of the library is a finding too
            p = [f"{type(e).__name__}: {e}"]
        if p is None:
            skipped += 1
        elif p:
            bad += 1
            print(f"seed {s}: " + " | ".join(p)[:1500], flush=True)
    print(f"{n} scenarios (seeds {first}..{first + n - 1}): {n - bad - skipped} identical, {bad} with differences, "
          f"{skipped} skipped (no defined reference answer: allocation list overflow" + (", or a frame without requests)" if ref_size else ")"))
    sys.exit(1 if bad else 0)
