"""The CPU oracle against the golden vectors of the reference's own unit tests (restated in tests/kat_cases.py)."""
import numpy as np
import pytest

from oracle.oracle import los_mask
from tests import kat_cases


@pytest.mark.parametrize('case', kat_cases.CASES, ids=lambda c: c.__name__)
def test_reference_known_answers(case):
    case('oracle')


def test_los_mask_symmetries_and_known_cells():
    """Row / column flips hold exactly; the transpose holds up to range 14.  At range >= 15 the reference's
    float64 rays hide (15,11) behind a blocker at (8,5) and (15,13) behind (8,6) -- exact rational arithmetic would
    not -- while the transposed pairs stay visible: the oracle reproduces exactly that."""
    R = 16
    for rd, cd in [(1, 0), (0, 2), (3, 1), (2, 5), (8, 5), (8, 6), (-4, 7), (6, -6)]:
        m = los_mask(R, rd, cd)
        np.testing.assert_array_equal(m[::-1], los_mask(R, -rd, cd))
        np.testing.assert_array_equal(m[:, ::-1], los_mask(R, rd, -cd))
        np.testing.assert_array_equal(m.T[2:-2, 2:-2], los_mask(R, cd, rd)[2:-2, 2:-2])    # offsets within +-14
        assert m[R + rd, R + cd] == 1                                   # a blocker never hides its own cell
    assert los_mask(R, 8, 5)[R + 15, R + 11] == 0 and los_mask(R, 5, 8)[R + 11, R + 15] == 1
    assert los_mask(R, 8, 6)[R + 15, R + 13] == 0 and los_mask(R, 6, 8)[R + 13, R + 15] == 1
    assert los_mask(R, 0, 0).all()                                       # a blocker on the viewer's cell hides nothing
    assert los_mask(2, 5, 5).all()                                       # out of range
    m = los_mask(3, 0, 1)                                                # directly behind a wall
    assert m[3, 5] == 0 and m[3, 6] == 0 and m[3, 4] == 1 and m[2, 4] == 1


def test_los_mask_equals_reference_fixture():
    """Every blocker offset at range 16, recorded from the unmodified reference's create_grid_and_mask
    (tests/golden/los_r16.npz, made by tests/golden/make_golden.py los)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'los_r16.npz'))
    R = int(g['range'])
    n = 2 * R + 1
    masks = np.unpackbits(g['packed'], axis=-1)[..., :n * n].reshape(n, n, n, n)
    for rd in range(-R, R + 1):
        for cd in range(-R, R + 1):
            np.testing.assert_array_equal(los_mask(R, rd, cd), masks[rd + R, cd + R], err_msg=f'blocker offset {(rd, cd)}')
