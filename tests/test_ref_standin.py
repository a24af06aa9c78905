"""The stand-in Eigen / Sophus headers (oracle/ref_standin, used only to build the reference's sources into oracle/_ref)
against numpy / scipy: inverses, SE3 exp, LLT, blocks, array() comparisons, row / segment views."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STANDIN = os.path.join(ROOT, "oracle", "ref_standin")


@pytest.fixture(scope="module")
def res(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("standin") / "selftest")
    subprocess.run(["/usr/bin/g++", "-std=c++14", "-O1", "-ffp-contract=off", "-I" + STANDIN, os.path.join(STANDIN, "test", "standin_selftest.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    return {k: np.array(v, np.float64) for k, v in json.loads(out).items()}


def test_inverses(res):
    K, T, G = res["K"], res["T"], res["G"]
    np.testing.assert_allclose(res["K_inv"] @ K, np.eye(4), atol=1e-6)
    np.testing.assert_allclose(res["T_inv"] @ T, np.eye(4), atol=1e-6)
    np.testing.assert_allclose(res["T_inv_sophus"], res["T_inv"], atol=1e-7)
    np.testing.assert_allclose(res["G_inv"], np.linalg.inv(G), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(res["KT"], K @ T, rtol=1e-6, atol=1e-4)


def test_se3_exp_is_the_matrix_exponential(res):
    from scipy.linalg import expm
    u, w = res["xi"][:3, 0], res["xi"][3:, 0]
    tw = np.zeros((4, 4))
    tw[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
    tw[:3, 3] = u
    np.testing.assert_allclose(res["T"], expm(tw), atol=2e-7)


def test_vector_ops(res):
    p, q, T = res["p"][:, 0], res["q"][:, 0], res["T"]
    np.testing.assert_allclose(res["Tp"][:, 0], T[:3, :3] @ p + T[:3, 3], atol=1e-6)
    np.testing.assert_allclose(res["hom"][:, 0], (T @ np.append(p, 1))[:3], atol=1e-6)
    np.testing.assert_allclose(res["cross"][:, 0], np.cross(p, q), atol=1e-6)
    np.testing.assert_allclose(res["normalized"][:, 0], p / np.linalg.norm(p), atol=1e-7)
    np.testing.assert_allclose(res["cwise"][:, 0], np.maximum(p * q, -0.3), atol=1e-7)
    np.testing.assert_array_equal(res["floor"][:, 0], np.floor(np.float32(p) * np.float32(1.7)))
    np.testing.assert_allclose(res["clamp"][:, 0], np.clip(p * 3, -1, 2), atol=1e-6)
    np.testing.assert_array_equal(res["cast"][:, 0], np.trunc(np.float32(p) * np.float32(2.6)))
    assert list(res["bools"]) == [1, 0, 0, 1]


def test_llt_rows_segments_blocks(res):
    np.testing.assert_allclose(res["llt_x"][:, 0], np.linalg.solve(res["A"], res["b"][:, 0]), rtol=1e-5)
    want = np.arange(8 * 32, dtype=np.float64).reshape(8, 32).sum(axis=0)[1:28]
    np.testing.assert_array_equal(res["rowsum_seg"][0], want)
    moved = res["T"].copy(); moved[:3, 3] += [1, 2, 3]
    np.testing.assert_allclose(res["block_add"], moved, atol=1e-6)
