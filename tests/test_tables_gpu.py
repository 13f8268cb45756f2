"""GPU: index tables are bit-exact with the reference (north_star: "Index permutations and bit-reversal tables must be
bit-exact")."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_bitrev_tables_bit_exact(gpu, orc, golden):
    for order in range(0, 11):
        assert np.array_equal(gpu.bitrev_table(order), golden[f"bitrev/{order}"])
    for order in (12, 16, 20):
        assert np.array_equal(gpu.bitrev_table(order), orc.bitrev_table(order))
    assert gpu.bitrev_table(4).tolist() == [0, 8, 4, 12, 2, 10, 6, 14, 1, 9, 5, 13, 3, 11, 7, 15]


def test_digitrev_bit_exact(gpu, orc, golden):
    for key in [k for k in golden.files if k.startswith("digitrev/")]:
        _, radix, size = key.split("/")
        assert np.array_equal(gpu.digitrev_perm(int(radix), int(size)), golden[key]), key
    for radix, size in ((4, 4096), (8, 4096), (3, 729), (5, 625)):
        assert np.array_equal(gpu.digitrev_perm(radix, size), orc.digitrev_perm(radix, size))
    with pytest.raises(RuntimeError):
        gpu.digitrev_perm(4, 24)


def test_fdl_index_sequence_bit_exact(gpu, orc, golden):
    # convolution/fdl_index_test.cpp:13-65
    wp, pairs = gpu.fdl_index_sequence(3, 6)
    assert wp.tolist() == [0, 1, 2, 0, 1, 2]
    assert pairs[0].tolist() == [[0, 0], [1, 2], [2, 1]]
    assert pairs[1].tolist() == [[0, 1], [1, 0], [2, 2]]
    for parts in (1, 2, 3, 4, 7):
        wp, pairs = gpu.fdl_index_sequence(parts, 2 * parts + 3)
        assert np.array_equal(wp, golden[f"fdl_index/{parts}/write_pos"])
        assert np.array_equal(pairs, golden[f"fdl_index/{parts}/pairs"])
    wp, pairs = gpu.fdl_index_sequence(1024, 5)
    owp, opairs = orc.fdl_index_sequence(1024, 5)
    assert np.array_equal(wp, owp) and np.array_equal(pairs, opairs)
