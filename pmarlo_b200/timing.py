"""CUDA-event stopwatch for pipeline stages and individual kernel launches."""

from __future__ import annotations

import contextlib
import os

import torch

# NVTX ranges per stage (tracing): on when PMB_NVTX=1, so that an nsys / ncu --nvtx
# timeline shows the stage names; off by default (a push/pop pair per stage costs ~1 us of host time).
_NVTX = os.environ.get("PMB_NVTX", "0") not in ("", "0")

__all__ = ["StageTimer", "NULL_TIMER"]


class StageTimer:
    """Events are recorded on torch's current stream, which is the stream every
    kernel of libpmb200 is launched on (``_lib.stream_handle``)."""

    def __init__(self, enabled: bool = True):
        self.enabled = enabled
        self.records: list[tuple[str, torch.cuda.Event, torch.cuda.Event]] = []

    @contextlib.contextmanager
    def stage(self, name: str):
        if _NVTX:
            torch.cuda.nvtx.range_push(f"pmb200:{name}")
        if not self.enabled:
            try:
                yield
            finally:
                if _NVTX:
                    torch.cuda.nvtx.range_pop()
            return
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        try:
            yield
        finally:
            b.record()
            self.records.append((name, a, b))
            if _NVTX:
                torch.cuda.nvtx.range_pop()

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
