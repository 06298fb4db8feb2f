"""Host-buffer front door for the multiscale RP-AdaIN transform.

The reference keeps features on the GPU, so its own call sites never cross PCIe; this module is the
"plugin call with HOST buffers" that `bench.py` times as `e2e`: per image it copies every level's
content / style / decoder-state planes from pinned host memory, runs the fused AdaIN(+blend) kernels
through the C ABI and copies the transformed levels back.  Copies and kernels are double-buffered on
three streams so PCIe (the bound) stays busy in both directions."""
from __future__ import annotations

from typing import List, Sequence

import torch

from . import functional as F


class MultiscaleHostPipe:
    """Reusable staging for `run(images)`: level l has `channels[l]` channels at `hw = h*w`;
    level order is shallow -> deep like `encode_rp_intermediate` (network/adain_rp.py:187-191)."""

    def __init__(self, channels: Sequence[int], h: int, w: int, device="cuda"):
        self.channels = list(channels)
        self.h, self.w = h, w
        self.device = torch.device(device)
        mk = lambda c: torch.empty(1, c, h, w, dtype=torch.float32, device=self.device)
        self.dev = [{"c": [mk(c) for c in channels], "s": [mk(c) for c in channels],
                     "p": [mk(c) for c in channels[:-1]], "o": [mk(c) for c in channels]} for _ in range(2)]
        self.s_in = torch.cuda.Stream(self.device)
        self.s_run = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self.h2d_bytes_per_image = sum(4 * c * h * w for c in channels) * 2 + sum(4 * c * h * w for c in channels[:-1])
        self.d2h_bytes_per_image = sum(4 * c * h * w for c in channels)

    def run(self, host_c: List[torch.Tensor], host_s: List[torch.Tensor], host_p: List[torch.Tensor],
            host_out: List[List[torch.Tensor]], images: int) -> None:
        """host_c/host_s/host_p: pinned per-level [1,C,H,W] tensors (re-used for every image: the
        bytes still cross PCIe each time); host_out: two sets of pinned per-level outputs."""
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_run = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]
        cur = torch.cuda.current_stream(self.device)
        for st in (self.s_in, self.s_run, self.s_out):
            st.wait_stream(cur)
        for i in range(images):
            b = i & 1
            d = self.dev[b]
            with torch.cuda.stream(self.s_in):
                if i >= 2:
                    self.s_in.wait_event(ev_run[b])       # device inputs of slot b are free again
                for l in range(len(self.channels)):
                    d["c"][l].copy_(host_c[l], non_blocking=True)
                    d["s"][l].copy_(host_s[l], non_blocking=True)
                    if l < len(self.channels) - 1:
                        d["p"][l].copy_(host_p[l], non_blocking=True)
                ev_in[b].record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(ev_in[b])
                if i >= 2:
                    self.s_run.wait_event(ev_out[b])      # device outputs of slot b were drained
                top = len(self.channels) - 1
                F._adain_raw(d["c"][top], d["s"][top], None, d["o"][top], self.channels[top] * self.h * self.w, F.EPS, False)
                for l in range(top - 1, -1, -1):
                    F._adain_raw(d["c"][l], d["s"][l], d["p"][l], d["o"][l], self.channels[l] * self.h * self.w, F.EPS, False)
                ev_run[b].record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_run[b])
                for l in range(len(self.channels)):
                    host_out[b][l].copy_(d["o"][l], non_blocking=True)
                ev_out[b].record(self.s_out)
        for st in (self.s_in, self.s_run, self.s_out):
            cur.wait_stream(st)
