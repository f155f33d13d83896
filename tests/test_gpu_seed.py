"""Seed candidate enumeration as a kernel against the REFERENCE'S CSeed::collectCandidates / unproject
(tests/golden/pmvs_seed.npz, from the reference objects): for every searched feature the candidate SET -- other image, other
feature, triangulated point, _response -- must be equal bit for bit, with every cell open and with a third of the cells closed
(CSeed::canAdd).  The order inside a feature's list is ascending _response (the reference sorts pointer values, see
cmvs-pmvs_b200/csrc/pmvs_seed.cuh)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def G(scene):
    g = np.load(os.path.join(HERE, "golden", "pmvs_seed.npz"))
    assert scene.sha256() == bytes(g["scene_sha256"]).hex()
    return g


@pytest.fixture(scope="module")
def seeded(gpu, scene, G):
    for i in range(scene.num):
        gpu.set_features(i, G["feat%d_xy" % i], G["feat%d_type" % i])
    return gpu


def _rows(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a[np.lexsort(a.T[::-1])]


@pytest.mark.parametrize("tag", ["open", "closed"])
def test_candidate_sets_equal_the_references(seeded, scene, G, tag):
    gpu = seeded
    cells = [gpu.grid_dims(i)[0] * gpu.grid_dims(i)[1] for i in range(scene.num)]
    blocked = np.zeros(sum(cells), np.uint8) if tag == "open" else (G["closed_counts"] >= 2).astype(np.uint8)
    for index in G["ref_images"]:
        out = gpu.seed_candidates(int(index), G["views%d" % index], blocked)
        rows, cand = G["%s_rows%d" % (tag, index)], G["%s_cand%d" % (tag, index)]
        # the same features are searched, in (cell, feature-in-cell) order
        assert np.array_equal(out["ref_cell"], rows[:, 0]), index
        assert np.array_equal(out["ref_count"], rows[:, 2]), index
        # ordinal of a feature inside its cell = its rank among the features of that cell
        ordinal = np.zeros(len(rows), np.int32)
        for k in range(1, len(rows)):
            ordinal[k] = ordinal[k - 1] + 1 if rows[k, 0] == rows[k - 1, 0] else 0
        assert np.array_equal(ordinal, rows[:, 1])
        pos = 0
        for k in range(len(rows)):
            n = rows[k, 2]
            s = out["ref_start"][k]
            oi = out["other_image"][s:s + n]; of = out["other_feature"][s:s + n]
            oxy = np.stack([G["feat%d_xy" % i][f] for i, f in zip(oi, of)]) if n else np.zeros((0, 2), np.float32)
            got = np.concatenate([oi[:, None].astype(np.float32), oxy, out["coords"][s:s + n], out["resp"][s:s + n, None]], axis=1) if n else np.zeros((0, 8), np.float32)
            want = cand[pos:pos + n]
            assert np.array_equal(_rows(got), _rows(want)), (index, k)
            assert (np.diff(out["resp"][s:s + n]) >= 0).all()
            pos += n
        assert pos == len(cand) and pos > 300


def test_no_views_and_all_blocked(seeded, scene, G):
    gpu = seeded
    cells = sum(gpu.grid_dims(i)[0] * gpu.grid_dims(i)[1] for i in range(scene.num))
    out = gpu.seed_candidates(0, np.zeros(0, np.int32), np.zeros(cells, np.uint8))
    assert len(out["ref_feature"]) == 0 and len(out["resp"]) == 0
    out = gpu.seed_candidates(0, G["views0"], np.ones(cells, np.uint8))
    assert len(out["ref_feature"]) == 0 and len(out["resp"]) == 0
