"""Stack files in the reference's on-disk formats (evcont_b200/io.py)."""
import numpy as np
import pytest

from conftest import synthetic_stack


@pytest.mark.parametrize("layout", [6, 5, 3, 2])
@pytest.mark.parametrize("index", [None, 3])
def test_stack_files_round_trip(tmp_path, layout, index):
    from evcont_b200 import io
    ovlp, one, two = synthetic_stack(4, 3, 2, layout)
    io.save_stack(tmp_path, ovlp, one, two, index=index)
    suffix = "" if index is None else f"_{index}"
    assert (tmp_path / f"two_rdm{suffix}.npy").exists()
    o2, one2, two2 = io.load_stack(tmp_path, index=index, mmap=(layout == 6))
    assert np.array_equal(o2, ovlp) and np.array_equal(one2, one) and np.array_equal(np.asarray(two2), two)
    assert two2.ndim == layout


def test_pair_directories_as_the_zundel_script_reads_them(tmp_path):
    from evcont_b200 import io
    n, N = 3, 4
    ovlp, one, two = synthetic_stack(n, N, 8, 2)
    io.save_pair_directories(tmp_path, ovlp, one, two)
    assert (tmp_path / "MPS_cross_2_1" / "two_rdm.npy").exists()
    o2, one2, two2 = io.load_pair_directories(tmp_path, N, n)
    assert np.array_equal(o2, ovlp) and np.array_equal(two2, two)
    il = np.tril_indices(N)
    assert np.array_equal(one2[il], one[il])
    # the mirror is the untransposed block, like the reference's assembly
    assert np.array_equal(one2[1, 2], one[2, 1])
