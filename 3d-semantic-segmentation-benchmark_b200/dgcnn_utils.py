"""Host-side mirror of the inference helper of the reference's models/dgcnn/utils.py on top of libpcnbr
(SURVEY.md 8f-4).

`predict_single_scene(model, points, device, batch_size, overlap)` keeps the reference's signature and results
(utils.py:67-131): a scene of N points is cut into windows of `batch_size` points every `batch_size - overlap` points,
every window goes through the model on its own (kNN graphs never cross a window), the logits are overlap-added and
averaged, and the prediction / confidence are the argmax / max softmax of the mean.  The reference runs the windows one
after the other at batch 1 and makes four passes over the (N,C) accumulators; here all full-length windows form ONE batch
(eval-mode BatchNorm keeps the windows independent, so the per-window logits are the same), the shorter windows at the end
of the scene form a second, zero-padded batch when the model has a length-aware forward (ours do), and one kernel
(`csrc/blocks.cu: window_merge_kernel`) does overlap-add, division, argmax and confidence.  CUDA only."""
from __future__ import annotations

import torch

from . import _lib
from .ops import _stream

__all__ = ["predict_single_scene", "scene_windows"]


def scene_windows(n_points: int, batch_size: int, overlap: int):
    """[(start, end)] of the reference's loop `for start in range(0, n, step)` (utils.py:108-110)."""
    step = batch_size - overlap
    if step <= 0:
        raise ValueError(f"overlap ({overlap}) must be smaller than the window ({batch_size})")
    return [(s, min(s + batch_size, n_points)) for s in range(0, n_points, step)]


def _logits(model, x, **kw):
    out = model(x, **kw)
    return out[0] if isinstance(out, tuple) else out


def _takes_lengths(model) -> bool:
    import inspect
    try:
        return "lengths" in inspect.signature(model.forward).parameters
    except (TypeError, ValueError):
        return False


def predict_single_scene(model, points: torch.Tensor, device: str = 'cuda', batch_size: int = 4096, overlap: int = 512,
                         max_windows_per_call: int = 64, return_logits: bool = False):
    """-> (predictions (N,) int64, confidences (N,) f32) on the CPU, as the reference (utils.py:67-131);
    `return_logits=True` appends the mean logits (N,C) (device tensor)."""
    if torch.device(device).type != "cuda":
        raise RuntimeError("pcnbr: predict_single_scene runs on a CUDA device (this build has no CPU fallback)")
    model.eval()
    points = points.to(device=device, dtype=torch.float32)
    n, F = points.shape
    with torch.no_grad():
        if n <= batch_size:                                           # utils.py:89-98
            windows, step = [(0, n)], max(n, 1)
            window = n
            parts = [_logits(model, points.T.unsqueeze(0)).reshape(n, -1)]
        else:
            windows, step, window = scene_windows(n, batch_size, overlap), batch_size - overlap, batch_size
            n_full = sum(1 for s, e in windows if e - s == batch_size)        # full windows come first
            full = points.unfold(0, batch_size, step)                        # (n_full', F, window) view, no copy
            parts = []
            for w0 in range(0, n_full, max_windows_per_call):
                x = full[w0:min(n_full, w0 + max_windows_per_call)]
                parts.append(_logits(model, x).reshape(x.shape[0] * batch_size, -1))
            tails = windows[n_full:]                                         # the shorter windows at the end of the scene
            if len(tails) > 1 and _takes_lengths(model):
                # length-aware models: the short windows go through the model as ONE zero-padded batch; the
                # padding takes no part in any graph, so every window's logits are those of the window passed alone
                xt = torch.zeros(len(tails), F, batch_size, dtype=torch.float32, device=points.device)
                for t, (s, e) in enumerate(tails):
                    xt[t, :, :e - s] = points[s:e].T
                parts.append(_logits(model, xt, lengths=[e - s for s, e in tails]).reshape(len(tails) * batch_size, -1))
                tail_rows = [batch_size] * len(tails)                        # rows each tail window occupies in `logits`
            else:
                for s, e in tails:
                    parts.append(_logits(model, points[s:e].T.unsqueeze(0)).reshape(e - s, -1))
                tail_rows = [e - s for s, e in tails]
        logits = (parts[0] if len(parts) == 1 else torch.cat(parts)).float().contiguous()
        C = logits.shape[1]
        offs, acc = [], 0
        for w, (s, e) in enumerate(windows):
            offs.append(acc)
            acc += (e - s) if (n <= batch_size or w < n_full) else tail_rows[w - n_full]
        win_off = torch.tensor(offs, dtype=torch.int64).to(points.device, non_blocking=True)
        mean = torch.empty(n, C, dtype=torch.float32, device=points.device) if return_logits else None
        pred = torch.empty(n, dtype=torch.int64, device=points.device)
        conf = torch.empty(n, dtype=torch.float32, device=points.device)
        _lib.call("pcnbr_window_merge_f32", logits.data_ptr(), win_off.data_ptr(), len(windows), n, window, step, C,
                  mean.data_ptr() if mean is not None else None, pred.data_ptr(), conf.data_ptr(), _stream())
    if return_logits:
        return pred.cpu(), conf.cpu(), mean
    return pred.cpu(), conf.cpu()
