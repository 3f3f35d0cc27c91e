"""``@callback(stage, interval)`` and the ``Callback`` base class (callback/callback.py:10-141 of the reference):
same trigger rules (int: every n-th step, float: every t seconds, callable: predicate)."""
from __future__ import annotations

from functools import wraps
from typing import Callable, Optional


def _validate_interval(interval) -> None:
    if not (isinstance(interval, (int, float)) or callable(interval)):
        raise TypeError(f"Invalid interval: {interval}. Must be int, float, or Callable")
    if isinstance(interval, float) and not (0 < interval < 1):
        raise ValueError(f"Invalid interval: {interval}. Must be between 0 and 1s if it is a float")
    if isinstance(interval, int) and not isinstance(interval, bool) and interval < 1:
        raise ValueError(f"Invalid interval: {interval}. Must be greater than 0 if it is an integer")


def _interval_triggered(sim, interval) -> bool:
    if callable(interval):
        return bool(interval(sim))
    if isinstance(interval, int):
        return sim.itime % interval == 0
    if isinstance(interval, float):
        return (sim.time % interval) < sim.dt
    return True


def callback(stage: Optional[str] = None, interval=1, needs_host: bool = True, reads=None, writes=None) -> Callable:
    """Extensions over the reference (mirror elision):
      needs_host=False   the callback uses only device-side diagnostics (sim.energies()); no mirror traffic at all;
      reads / writes     names of what the callback reads / modifies on the host mirrors -- field attributes ("ex" .. "rho"),
                         "fields", "psi", "particles".  Only those cross PCIe before / after the stage; a read-only
                         diagnostic (writes=()) costs no upload.  Without hints everything is synced both ways."""
    def decorator(func: Callable) -> Callable:
        _validate_interval(interval)

        @wraps(func)
        def wrapper(*args, **kwargs):
            sim = args[-1]
            if not _interval_triggered(sim, interval):
                return None
            ret = func(*args, **kwargs)
            sim.mpi.comm.Barrier()
            return ret
        wrapper.stage = stage
        wrapper.interval = interval
        wrapper.needs_host = needs_host  # False: uses only device-side diagnostics, mirrors are not synced for it
        wrapper.reads = None if reads is None else tuple(reads)
        wrapper.writes = None if writes is None else tuple(writes)
        return wrapper
    return decorator


class Callback:
    interval = 1
    stage = None
    needs_host = True
    reads = None   # see callback(): names touched on the host mirrors, None = everything
    writes = None

    def __call__(self, sim):
        _validate_interval(self.interval)
        if not _interval_triggered(sim, self.interval):
            return None
        ret = self._call(sim)
        sim.mpi.comm.Barrier()
        return ret

    def _call(self, sim):
        raise NotImplementedError
