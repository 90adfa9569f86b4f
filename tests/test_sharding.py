"""Work-list sharding: pcdms_b200.sharding.split_list_into_chunks against the reference's own helper,
extracted from the driver source at test time (skipped where /root/reference is absent), and its defining properties."""
import ast
from pathlib import Path

import pytest

from pcdms_b200.sharding import rank_shard, split_list_into_chunks

REF = Path("/root/reference")
DRIVERS = ["stage1_batchtest_prior_model.py", "stage2_batchtest_inpaint_model.py", "stage3_batchtest_refined_model.py"]


def _reference_fn(driver):
    src = (REF / driver).read_text()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "split_list_into_chunks")
    ns = {}
    exec(compile(ast.Module(body=[node], type_ignores=[]), driver, "exec"), ns)
    return ns["split_list_into_chunks"]


@pytest.mark.skipif(not REF.exists(), reason="/root/reference not present")
@pytest.mark.parametrize("driver", DRIVERS)
def test_matches_the_reference_helper(driver):
    ref = _reference_fn(driver)
    for length in list(range(1, 40)) + [64, 100, 8570]:
        for n in (1, 2, 3, 4, 7, 8):
            lst = list(range(length))
            if length < n:
                with pytest.raises(ValueError):
                    ref(list(lst), n)
                with pytest.raises(ValueError):
                    split_list_into_chunks(lst, n)
            else:
                assert split_list_into_chunks(lst, n) == ref(list(lst), n), (length, n)


def test_partition_properties():
    for length in (8, 9, 16, 64, 1001, 8570):
        for n in (1, 2, 4, 8):
            lst = [f"pair{i}" for i in range(length)]
            chunks = split_list_into_chunks(lst, n)
            assert [x for c in chunks for x in c] == lst                       # a partition, order preserved
            if length % n < max(length // n, 1) or n == 1:                     # the usual case: exactly n chunks,
                assert len(chunks) == n                                        # equal but for the remainder on the last
                assert all(len(c) == length // n for c in chunks[:-1])
                assert len(chunks[-1]) == length // n + length % n
            assert all(rank_shard(lst, r, n) == chunks[r] for r in range(n))
    assert split_list_into_chunks(list(range(64)), 8)[3] == list(range(24, 32))   # BASELINE config 4: 8 per rank
    assert len(split_list_into_chunks(list(range(5)), 3)) == 4                    # the reference's short-list quirk
    with pytest.raises(ValueError):
        rank_shard([1, 2, 3], 3, 3)
    with pytest.raises(ValueError):
        split_list_into_chunks([1, 2], 0)
