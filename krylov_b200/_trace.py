"""Optional stage timer (KRYLOV_B200_TRACE=1): ``mark(label)`` synchronises the device and
records a timestamp; ``take()`` returns [(label, seconds since the previous mark)].  Off by
default -- a mark then costs one dict lookup."""
import os
import time

_ON = os.environ.get("KRYLOV_B200_TRACE", "") not in ("", "0")
_T = []


def enable(on=True):
    global _ON
    _ON = bool(on)
    _T.clear()


def mark(label):
    if not _ON:
        return
    import torch

    torch.cuda.synchronize()
    _T.append((label, time.perf_counter()))


def take():
    out = [(b[0], b[1] - a[1]) for a, b in zip(_T, _T[1:])]
    _T.clear()
    return out
