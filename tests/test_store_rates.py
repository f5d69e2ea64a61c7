"""Compaction of the rates map into the `_rates` records (SURVEY.md section
8f row 3, the `_store` half): oqg_rates_to_records against the reference's
MapArray.to_array and the chunk grouping of _store
(oracle/make_golden_store.py -> tests/golden/store_rates.npz).  Integer and
byte work: bit-exact, record by record."""
import os
import numpy as np
import pytest

from helpers import GOLDEN

D = np.load(os.path.join(GOLDEN, 'store_rates.npz'))


def test_fixture_is_what_the_docstring_says():
    recs = D['records']
    assert recs.dtype.names == ('sid', 'lid', 'gid', 'rate')
    assert recs.dtype.itemsize == 12
    assert (recs['rate'] > 0).all() and len(recs) == (D['array'] > 0).sum()


@pytest.mark.gpu
@pytest.mark.parametrize('device_map', [False, True])
def test_rates_to_records(device_map):
    import oq_engine_b200 as oq
    oq._lib.lib()
    N, L, G = D['array'].shape
    rmap = oq.MapArray(D['sids'], L, G, rates=True,
                       device='cuda:0' if device_map else None).fill(0)
    if device_map:
        import torch
        rmap.array = torch.from_numpy(D['array']).to('cuda:0')
    else:
        rmap.array = D['array'].copy()
    recs = rmap.to_array(D['gid'])
    assert recs.dtype == oq.rates_dt
    assert recs.tobytes() == D['records'].tobytes()
    for nc in (1, 4, 7):
        got, slices = oq.rates_to_records(rmap, D['gid'], num_chunks=nc)
        assert got.tobytes() == D['records_%d' % nc].tobytes()
        np.testing.assert_array_equal(
            np.stack([slices['idx'], slices['start'], slices['stop']], 1),
            D['slices_%d' % nc])
    # an empty map gives no record
    rmap0 = oq.MapArray(D['sids'], L, G, rates=True).fill(0)
    recs0, sl0 = oq.rates_to_records(rmap0, D['gid'], num_chunks=4)
    assert len(recs0) == 0 and (sl0['start'] == sl0['stop']).all()
    # level offset and a level index beyond 16 bits
    recs1, _ = oq.rates_to_records(rmap, D['gid'], level0=100)
    np.testing.assert_array_equal(recs1['lid'], D['records']['lid'] + 100)
    with pytest.raises(RuntimeError, match='16 bits'):
        oq.rates_to_records(rmap, D['gid'], level0=65500)
