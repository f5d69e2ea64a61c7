"""Monitors of the hot path, with the names the reference's calculators read
(`openquake/baselib/performance.py:222-330`, Monitor; `ContextMaker.
init_monitoring`, `openquake/hazardlib/contexts.py:681-690`): a Monitor is a
context manager accumulating `duration` (seconds) and `counts`, `monitor(op)`
makes a child.  Device work is asynchronous: a monitor created with
``device_sync=True`` (inherited by its children) synchronises the device when
it is entered and left, so that its duration is the time of the kernels
launched inside it; by default nothing synchronises and a duration is the
host time spent enqueuing.
"""
import time


class Monitor(object):
    """
    >>> mon = Monitor('classical')
    >>> child = mon('get_poes')
    >>> with child: pass
    >>> child.counts
    1
    """
    def __init__(self, operation='', measuremem=False, inner_loop=False,
                 device_sync=False):
        self.operation = operation
        self.measuremem = measuremem      # accepted, host memory not sampled
        self.inner_loop = inner_loop
        self.device_sync = device_sync
        self.duration = 0.
        self.counts = 0
        self.children = []
        self.task_no = -1
        self._t0 = None

    def _sync(self):
        if self.device_sync:
            import torch
            if torch.cuda.is_available():
                torch.cuda.synchronize()

    def __call__(self, operation='no operation', measuremem=False, **kw):
        child = Monitor(operation, measuremem,
                        kw.get('inner_loop', False), self.device_sync)
        child.task_no = self.task_no
        self.children.append(child)
        return child

    def __enter__(self):
        self._sync()
        self._t0 = time.time()
        return self

    def __exit__(self, etype, exc, tb):
        self._sync()
        self.duration += time.time() - self._t0
        self.counts += 1

    @property
    def dt(self):
        """Last time interval measured, if the monitor is running"""
        return time.time() - self._t0

    def get_data(self):
        """[(operation, time_sec, counts)] of this monitor and its children
        (the columns of the reference's performance_data that apply)"""
        data = []
        if self.counts:
            data.append((self.operation, self.duration, self.counts))
        for child in self.children:
            data.extend(child.get_data())
        return data

    def reset(self):
        self.duration = 0.
        self.counts = 0

    def __repr__(self):
        return '<%s %s, duration=%ss, counts=%d>' % (
            self.__class__.__name__, self.operation, self.duration,
            self.counts)
