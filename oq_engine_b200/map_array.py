"""MapArray: per-site probabilities of no exceedance or rates, mirroring
openquake/hazardlib/map_array.py:287-335, 443-455, 511-563.

``array`` is float32 of shape (N, L, G) exactly as in the reference.  It
lives on the host (numpy) by default; ``MapArray(..., device='cuda:0')`` keeps
it on the GPU as a torch tensor so that consecutive updates do not cross
PCIe.  The updates themselves run in CUDA (ContextMaker.update /
update_indep -> oqg_update / oqg_hazard_planar).
"""
import numpy

F32 = numpy.float32


class MapArray(object):
    def __init__(self, sids, shape_y, shape_z, rates=False, device=None):
        self.sids = numpy.asarray(sids, numpy.uint32)
        self.shape = (len(self.sids), shape_y, shape_z)
        self.rates = rates
        self.device = device
        self.array = None

    @property
    def sidx(self):
        """site ID -> row index (map_array.py:296-304)"""
        if not hasattr(self, '_sidx'):
            idxs = numpy.zeros(int(self.sids.max()) + 1 if len(self.sids)
                               else 0, numpy.uint32)
            idxs[self.sids] = numpy.arange(len(self.sids), dtype=numpy.uint32)
            self._sidx = idxs
        return self._sidx

    def new(self, array):
        new = MapArray(self.sids, array.shape[1], array.shape[2], self.rates,
                       self.device)
        new.array = array
        return new

    def fill(self, value):
        assert 0 <= value <= 1, value
        if self.device is None:
            self.array = numpy.empty(self.shape, F32)
            self.array.fill(value)
        else:
            import torch
            self.array = torch.full(self.shape, float(value),
                                    dtype=torch.float32, device=self.device)
        return self

    def numpy(self):
        """The array on the host"""
        if self.device is None:
            return self.array
        return self.array.cpu().numpy()

    def update_indep(self, poes, ctxt, itime, rates=None):
        """Update probabilities for independent ruptures
        (map_array.py:443-455); poes (n, L, G)"""
        from .contexts import update_device
        if rates is None:
            rates = ctxt.occurrence_rate
        update_device(self, poes, rates, ctxt.probs_occur,
                      self.sidx[ctxt.sids], itime)

    def to_rates(self, itime=1.):
        """map_array.py:511-525: pnes -> rates"""
        if self.rates:
            return self
        pnes = self.array
        # map_array.py:386-394: a PoE of exactly 1 becomes 1 - 1.11E-16
        if self.device is None:
            pnes = numpy.where(pnes == 0., F32(1.11E-16), pnes)
            rates = - numpy.log(pnes)
        else:
            import torch
            pnes = torch.where(pnes == 0., torch.full_like(pnes, 1.11E-16),
                               pnes)
            rates = - torch.log(pnes)
        if itime != 1.:
            rates = rates / itime
        new = self.new(rates)
        new.rates = True
        return new

    def to_array(self, gid):
        """map_array.py:396-405: the rates > 0 as records (sid, lid, gid,
        rate), ordered by gsim column, site and level (device compaction)"""
        from .postproc import rates_to_records
        return rates_to_records(self, gid)[0]

    def __invert__(self):
        return self.new(1. - self.array)

    def __iadd__(self, other):
        # RateMap.__iadd__, map_array.py:558-563 (same sites)
        assert self.rates and other.rates
        self.array += other.array
        return self

    def __repr__(self):
        return '<MapArray(%d, %d, %d)%s>' % (
            self.shape + (' rates' if self.rates else '',))
