"""Optional per-stage CUDA-event timing of the hot path (used by bench.py for the roofline and
the stage breakdown).  Inactive by default: `stage()` is then a no-op context manager."""
import contextlib

import torch

_active = None


class StageTimer:
    def __init__(self):
        self.events = {}      # name -> list of (start, end)

    def __enter__(self):
        global _active
        _active = self
        return self

    def __exit__(self, *exc):
        global _active
        _active = None

    def totals_ms(self):
        """Call after torch.cuda.synchronize().  name -> (total ms, number of timed calls)."""
        return {k: (sum(s.elapsed_time(e) for s, e in v), len(v)) for k, v in self.events.items()}


@contextlib.contextmanager
def stage(name):
    t = _active
    if t is None:
        yield
        return
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    try:
        yield
    finally:
        e.record()
        t.events.setdefault(name, []).append((s, e))
