"""GPU tests of the single-process multi-device driver (mcd_cell2cell_multi) behind ``MaCroDNA(..., devices=[...])``.
The test box has one GPU, so the device list names it several times: separate contexts, separate workspaces, peer
copies that degenerate to device-to-device copies -- every code path of the sharded driver except a physical NVLink
hop.  (bench.py asserts N-GPU == 1-GPU bit for bit on the real multi-GPU box.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["ozaki", "fp64", "split"])
@pytest.mark.parametrize("ndev", [2, 3])
def test_sharded_driver_equals_single_device(precision, ndev):
    from macrodna_b200 import MaCroDNA, synth

    inst = synth.make_arrays(1001, 140, 700, 4, seed=23)  # 1001 rows: uneven shards; 8 steps
    rna_df, dna_df, lab = synth.make_frames(inst)          # shuffled genes + extra RNA genes: device gene gather
    one = MaCroDNA(rna_df, dna_df, lab, precision=precision)
    r1, t1 = one.cell2cell_assignment()
    many = MaCroDNA(rna_df, dna_df, lab, precision=precision, devices=[0] * ndev)
    r2, t2 = many.cell2cell_assignment()
    assert (one.last_assign == many.last_assign).all() and (one.last_step == many.last_step).all()
    assert (one.last_objective == many.last_objective).all()
    assert r1.equals(r2) and t1.equals(t2)
    assert many.last_stats["cert_rel_gap"] <= 1e-12 and many.last_stats["cert_steps"] == 8
    # the matrix is resident on the first device: views work after a sharded run
    sub = many.subinstance_assignment(dna_cells=list(dna_df.columns[:100]))
    ref = one.subinstance_assignment(dna_cells=list(dna_df.columns[:100]))
    assert sub.equals(ref)
    clone = MaCroDNA(rna_df, dna_df, lab, devices=[0, 0]).cell2clone_assignment()
    assert list(clone.columns) == ["predict_cell", "predict_clone"]


def test_sharded_driver_more_devices_than_rows_and_nan():
    from macrodna_b200 import _lib, get_handle
    from macrodna_b200.api import get_handles

    rng = np.random.default_rng(1)
    rna, dna = rng.random((3, 40)), rng.random((5, 40))
    hs = get_handles([0, 0, 0, 0])               # 4 shards for 3 RNA rows: one shard is empty
    a, s, o, st = _lib.cell2cell_multi(hs, rna, dna, 3, 5, 40)
    a1, s1, o1, _ = get_handle(0).cell2cell(rna, dna, 3, 5, 40)
    assert (a == a1).all() and (s == s1).all() and (o == o1).all()
    rna[2, 7] = np.inf                            # seen by the LAST non-empty shard only
    with pytest.raises(ValueError, match="non-finite|NaN"):
        _lib.cell2cell_multi(hs, rna, dna, 3, 5, 40)
    rna[2, 7] = 0.5
    a2, _, _, _ = _lib.cell2cell_multi(hs, rna, dna, 3, 5, 40)   # handles are usable afterwards
    assert a2.shape == (3,)
