"""The CUDA engine (through the C-ABI) against the golden vectors of the reference's own unit tests."""
import pytest

from tests import kat_cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('case', kat_cases.CASES, ids=lambda c: c.__name__)
def test_reference_known_answers_on_the_engine(case):
    case('engine')
