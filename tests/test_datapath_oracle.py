"""CPU: the integer / data-path oracle (patch indexing, OOB extraction, float64 stitch, label
assembly) against the reference's own functions (tests/golden/datapath.npz) and the host-side
product code of nas_3d_unet_b200.infer - all bit-exact."""
import numpy as np
import pytest

from conftest import GOLDEN
from oracle import nas3d_oracle as O
import os


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(GOLDEN, "datapath.npz"))


def _cases(G):
    i = 0
    while 'patching/%d/args' % i in G.files:
        a = G['patching/%d/args' % i]
        yield tuple(a[:3]), tuple(a[3:6]), (None if a[6] < 0 else int(a[6])), G['patching/%d/corners' % i]
        i += 1


def test_patching_matches_reference(G):
    from nas_3d_unet_b200 import infer
    n = 0
    for img, ps, ov, ref in _cases(G):
        np.testing.assert_array_equal(O.patching(img, ps, overlap=ov), ref)
        np.testing.assert_array_equal(infer.patching(img, ps, overlap=ov), ref)
        n += 1
    assert n == 8
    # SURVEY App. D known answers
    c = O.patching((240, 240, 155), (128, 128, 128))
    assert len(c) == 9 and tuple(c[-1]) == (56, 56, 13) and set(c[:-1, 0]) == {0, 112} and set(c[:-1, 2]) == {0, 27}


def test_get_patch_oob_matches_reference(G):
    data = np.arange(4 * 10 ** 3, dtype=np.float32).reshape(4, 10, 10, 10)
    p = O.get_patch(data, (8, 8, 8), (-3, 5, 2))
    np.testing.assert_array_equal(p, G['get_patch/kat'])
    assert list(p[0, :, 0, 0]) == [0, 0, 0, 52, 152, 252, 352, 452]


def test_stitch_matches_reference_bit_exact(G):
    patches = G['stitch/patches']
    corners = G['stitch/corners']
    st = O.stitch(list(patches), corners, G['stitch/result'].shape)
    assert st.dtype == np.float64
    np.testing.assert_array_equal(st, G['stitch/result'])


def test_label_assembly_matches_reference(G):
    p = G['tumor/pred']
    np.testing.assert_array_equal(O.tumor_pred(p, 0.5, True), G['tumor/inclusive'])
    np.testing.assert_array_equal(O.tumor_pred(p, 0.5, False), G['tumor/exclusive'])
    t = G['labels/truth']
    np.testing.assert_array_equal(O.multi_class_labels(t, True), G['labels/inclusive'])
    np.testing.assert_array_equal(O.multi_class_labels(t, False), G['labels/exclusive'])
    # the logical_or quirk: label 4 is not part of channel 1 (WT)
    tt = np.array([0, 1, 2, 4, 3], dtype=np.int16).reshape(1, 1, 5, 1, 1)
    y = O.multi_class_labels(tt, True)[0, :, :, 0, 0]
    assert y.tolist() == [[0, 1, 0, 1, 0], [0, 1, 1, 0, 0], [0, 0, 0, 1, 0]]


def test_permutation_matches_reference_all_48_keys():
    """augment.permute_data (augment.py:105-132): oracle restatement and the host index map of the
    CUDA staging kernel against the reference's outputs for every key (tests/golden/permute.npz)"""
    import os
    from conftest import ROOT
    from nas_3d_unet_b200.data import index_map, permutation_keys
    z = np.load(os.path.join(ROOT, "tests", "golden", "permute.npz"))
    keys = O.permutation_keys()
    assert keys == permutation_keys() and len(keys) == 48
    assert [[k[0][0], k[0][1], k[1], k[2], k[3], k[4]] for k in keys] == z["keys"].tolist()
    data = z["data"]
    i, j, l = np.meshgrid(range(5), range(5), range(5), indexing="ij")
    for n, k in enumerate(keys):
        np.testing.assert_array_equal(O.permute_data(data, k), z["out"][n])
        b, s0, s1, s2 = index_map(k, (5, 5, 5))
        np.testing.assert_array_equal(data.reshape(2, -1)[:, b + s0 * i + s1 * j + s2 * l], z["out"][n])
    # SURVEY App. D KAT
    kat = O.permute_data(np.arange(54).reshape(2, 3, 3, 3), ((0, 1), 0, 1, 0, 1))[0, 0]
    assert kat.tolist() == [[0, 9, 18], [1, 10, 19], [2, 11, 20]]
    # flips alone are defined for non-cubic patches; rotations are not (generator.py:213)
    assert index_map(((0, 0), 1, 0, 1, 0), (4, 6, 8)) == (151, -48, 8, -1)
    with pytest.raises(AssertionError):
        index_map(((0, 1), 0, 0, 0, 0), (4, 6, 8))
