"""CUDA-event stopwatch for pipeline stages and individual kernel launches."""

from __future__ import annotations

import contextlib

import torch

__all__ = ["StageTimer", "NULL_TIMER"]


class StageTimer:
    """Events are recorded on torch's current stream, which is the stream every
    kernel of libpmb200 is launched on (``_lib.stream_handle``)."""

    def __init__(self, enabled: bool = True):
        self.enabled = enabled
        self.records: list[tuple[str, torch.cuda.Event, torch.cuda.Event]] = []

    @contextlib.contextmanager
    def stage(self, name: str):
        if not self.enabled:
            yield
            return
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        try:
            yield
        finally:
            b.record()
            self.records.append((name, a, b))

    def totals_ms(self) -> dict[str, float]:
        out: dict[str, float] = {}
        for name, a, b in self.records:
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out

    def counts(self) -> dict[str, int]:
        out: dict[str, int] = {}
        for name, _, _ in self.records:
            out[name] = out.get(name, 0) + 1
        return out

    def reset(self) -> None:
        self.records.clear()


NULL_TIMER = StageTimer(False)
