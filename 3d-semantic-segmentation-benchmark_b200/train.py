"""Training-step plumbing around the hot path: the masked one-hot cross entropy the reference's
train.py uses, and a batch-sharded data-parallel gradient exchange (one flat fp32 bucket, one NCCL
all-reduce per step over NVLink).  The reference itself is single-device; per-cloud
ops are independent, so the only collective is the gradient sum.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn.functional as F


class _MaskedCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, onehot, lengths):
        from . import _lib
        from .ops import _stream
        B, L, C = logits.shape
        dev = logits.device
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        need = ctx.needs_input_grad[0]
        grad = torch.empty_like(logits) if need else None
        partial = torch.empty(B * _lib.size("pcnbr_masked_ce_blocks", L), dtype=torch.float32, device=dev)
        _lib.call("pcnbr_masked_ce_f32", logits.data_ptr(), onehot.data_ptr(), lengths.data_ptr(), B, L, C, loss.data_ptr(),
                  grad.data_ptr() if need else None, partial.data_ptr(), _stream())
        if need:
            ctx.save_for_backward(grad)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None


def masked_onehot_cross_entropy(logits: torch.Tensor, targets_onehot: torch.Tensor, pad_starts: torch.Tensor):
    """Mean cross entropy over the unpadded points; same value as the reference's Training/train_model.py:15-57 but
    without its host sync (`total_non_pad.item()`, :53).  On CUDA fp32 logits with <= 64 classes: one fused kernel that
    also produces the gradient (csrc/loss.cu); otherwise the reference's op sequence in torch."""
    B, L, C = logits.shape
    if logits.is_cuda and logits.dtype == torch.float32 and C <= 64:
        onehot = targets_onehot.to(device=logits.device, dtype=torch.uint8).contiguous()
        lens = pad_starts.to(device=logits.device, dtype=torch.int64).contiguous()
        return _MaskedCEFn.apply(logits.contiguous(), onehot, lens)
    from .ops import note_fallback
    note_fallback("masked_onehot_cross_entropy: torch ops")
    logp = F.log_softmax(logits, dim=-1)
    tok = -(targets_onehot.to(logp.dtype) * logp).sum(dim=-1)
    mask = (torch.arange(L, device=logits.device).unsqueeze(0) < pad_starts.to(logits.device).long().unsqueeze(1)).to(logp.dtype)
    return (tok * mask).sum() / mask.sum().clamp_min(1.0)


def evaluate(model: torch.nn.Module, test_loader, criterion=None, device: str = "cuda", length_aware: bool = False):
    """The reference's validation loop (Training/training.py:80-133) with the same return tuple
    (mean loss, accuracy, mean IoU, per-class IoU (C,), confusion matrix (C,C) int64) -- but the whole validation set
    accumulates into ONE device-side confusion matrix and one device-side loss sum, read once at the end, instead of
    B x C (x C) `.item()` round trips per batch (Training/metrics.py:97-110).
    length_aware=True passes the loader's `lengths` to the model (forward(points, lengths=...)): the zero
    padding of collate_blocks then takes no part in FPS / grouping / kNN, i.e. every cloud is evaluated as if it had been
    passed alone; False reproduces the reference exactly (the padding participates, training.py:112)."""
    from . import metrics
    criterion = criterion or masked_onehot_cross_entropy
    model.eval()
    matrix, loss_sum, batches = None, None, 0
    with torch.no_grad():
        for points, labels, lengths in test_loader:
            points = points.to(device)
            labels = labels.to(device)
            outputs = model(points, lengths=lengths) if length_aware else model(points)
            outputs = outputs[0] if isinstance(outputs, tuple) else outputs
            loss = criterion(outputs, labels, lengths.long()).detach().float()
            loss_sum = loss if loss_sum is None else loss_sum + loss
            if matrix is None:
                matrix = torch.zeros(outputs.shape[-1], outputs.shape[-1], dtype=torch.int64, device=outputs.device)
                unl = torch.zeros(outputs.shape[-1], dtype=torch.int64, device=outputs.device)
            metrics.confusion_matrix_device(outputs, labels, lengths, out=matrix, unlabeled=unl)   # argmax of logits == argmax of softmax
            batches += 1
    if batches == 0:
        raise ValueError("evaluate: empty loader")
    m = matrix.cpu()
    inter, union = (t.to(torch.float32) for t in metrics._iou_terms(m, unl.cpu()))
    eps = 1e-6
    ious = (inter + eps) / (union + eps)
    return (loss_sum / batches).item(), (m.diagonal().sum() / m.sum()).item(), ious.mean().item(), ious, m


class FlatGradBucket:
    """All parameter gradients live in ONE contiguous fp32 buffer (p.grad are views into it), so the data-parallel
    exchange moves 4-17 MB in at most two all-reduces instead of 90-150 small ones, and zeroing the gradients is one memset.

    Overlap (world > 1): the buffer is cut into an EARLY part -- the parameters whose gradients the backward pass produces
    first, i.e. the ones registered last: the network head, >= 90 % of the bytes -- and a LATE part (the first layers).
    A post-accumulate-grad hook counts the early gradients in; when the last of them has landed, the early part is packed,
    pre-scaled by 1/world and its all-reduce is started asynchronously (NCCL's own stream), so it runs under the backward of
    the first layers -- the EdgeConv chain of DGCNN, sa2 / sa1 of PointNet++: most of the backward's time.
    all_reduce_mean() then only exchanges the small late part and waits for the early one.  Inside a captured CUDA graph
    the side stream forks from and re-joins the capturing stream through the events NCCL's Work objects record."""

    def __init__(self, module: torch.nn.Module, group=None, steal_grads: bool = False, overlap: bool = True,
                 late_fraction: float = 0.10):
        """steal_grads: zero() drops the gradients instead of clearing the bucket, so autograd hands each freshly
        computed gradient tensor to its parameter (no accumulate kernel per parameter: ~90 tiny launches per PointNet++
        step).  With several ranks the gradients are packed into the bucket with one multi-tensor copy per part, reduced,
        and every p.grad is pointed at its (reduced) bucket view; with one rank there is nothing to do.
        late_fraction: upper bound on the share of the bucket (in elements) left to the late part."""
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.group = group
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views, self.offsets = [], []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            self.offsets.append(off)
            off += p.numel()
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.steal = bool(steal_grads)
        if not self.steal:
            for p, v in zip(self.params, self.views):
                p.grad = v
        # ---- early / late split (registration order = forward order; the backward produces gradients in reverse)
        self.split = 0                                       # params[:split] = late part, params[split:] = early part
        self.overlap = bool(overlap) and self.world > 1 and len(self.params) > 1
        self._pending = None                                 # Work of the early all-reduce in flight
        self._arrived = 0
        self._hooks = []
        if self.overlap:
            acc = 0
            for i, p in enumerate(self.params):
                if acc + p.numel() > late_fraction * n:
                    break
                acc += p.numel()
                self.split = i + 1
            if self.split == 0 or self.split == len(self.params):
                self.overlap = False
        if self.overlap:
            for p in self.params[self.split:]:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_early_grad))

    # ---- helpers
    def _part(self, lo: int, hi: int) -> torch.Tensor:
        start = self.offsets[lo]
        end = self.offsets[hi] if hi < len(self.params) else self.flat.numel()
        return self.flat[start:end]

    def _pack(self, lo: int, hi: int) -> None:
        """Bring params[lo:hi]'s gradients into their bucket views (steal mode) and pre-scale the part by 1/world."""
        part = self._part(lo, hi)
        if self.steal:
            have = [(v, p.grad) for p, v in zip(self.params[lo:hi], self.views[lo:hi]) if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
            if len(have) != hi - lo:
                missing = [v for p, v in zip(self.params[lo:hi], self.views[lo:hi]) if p.grad is None]
                for v in missing:
                    v.zero_()                                 # parameters that received no gradient contribute 0
            if have:
                torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        part.mul_(1.0 / self.world)

    def _on_early_grad(self, _param) -> None:
        self._arrived += 1
        if self._arrived == len(self.params) - self.split and self._pending is None:
            self._pack(self.split, len(self.params))
            self._pending = dist.all_reduce(self._part(self.split, len(self.params)), op=dist.ReduceOp.SUM, group=self.group,
                                            async_op=True)

    def zero(self) -> None:
        self._arrived = 0
        if self.steal:
            for p in self.params:
                p.grad = None
            return
        self.flat.zero_()

    def all_reduce_mean(self) -> None:
        """Sum the bucket over ranks and divide by the world size (standard DDP semantics).  With overlap the early part is
        normally already in flight (started by the hook during the backward): only the late part is exchanged here."""
        if self.world <= 1:
            return
        if self.overlap and self._pending is not None:
            self._pack(0, self.split)
            dist.all_reduce(self._part(0, self.split), op=dist.ReduceOp.SUM, group=self.group)
            self._pending.wait()
            self._pending = None
        else:                                                 # no overlap (or a backward that skipped early parameters)
            self._pack(0, len(self.params))
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self._arrived = 0
        if self.steal:
            for p, v in zip(self.params, self.views):
                p.grad = v


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank start from rank `src`'s parameters and buffers."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def shard_batch(n_clouds: int, rank: int, world: int) -> slice:
    """Contiguous slice of the global batch owned by `rank` (clouds are independent units)."""
    per = (n_clouds + world - 1) // world
    return slice(min(rank * per, n_clouds), min((rank + 1) * per, n_clouds))


class GraphedTrainStep:
    """One full train step (forward, loss, backward, gradient all-reduce, Adam) captured ONCE into a CUDA graph and
    replayed per batch: ~220 libpcnbr + ~400 library launches per step collapse into one graph launch, so the step
    is bounded by the GPU, not by Python/launch overhead (CPU issue time was 15.6 ms of a 17 ms step).

    Every libpcnbr entry point is capture-safe (asynchronous on the given stream, no host sync, workspaces come from
    torch's graph-private pool).  Inputs are copied into static buffers; the loss comes back as a static tensor.

    geometry_fn (optional): `geometry_fn(model, *batch, stream=...)` -> list of tensors (+ trailing keep-alive objects) and
    `forward_fn(model, *batch, geometry=tensors)` consumes them (PointNetpp / DGCNN.prepare_geometry: FPS picks, ball-query /
    kNN tables and their CSR inverses depend on the input coordinates only).  The step is then software-pipelined over
    two batches: a call hands in batch i+1, whose geometry is computed on the side stream while the captured step trains on
    batch i (handed in by the previous call) -- FPS of a 4096-point cloud occupies one SM per cloud for 0.6 ms and no
    longer sits at the head of the step.  __call__ therefore returns the loss of the PREVIOUS call's batch (the first
    call trains on the example batch given at construction); flush() trains on the batch still in the pipeline."""

    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, bucket: FlatGradBucket,
                 forward_fn, example_batch, warmup: int = 3, geometry_fn=None):
        self.model, self.opt, self.bucket, self.forward_fn = model, optimizer, bucket, forward_fn
        self.geometry_fn = geometry_fn
        self.static = [t.clone() for t in example_batch]
        if geometry_fn is not None:
            self.static_next = [t.clone() for t in example_batch]
            first = [t for t in geometry_fn(model, *self.static_next) if isinstance(t, torch.Tensor)]
            self.geo_next = [t.clone() for t in first]               # geometry of static_next (ready before a replay)
            self.geo_cur = [t.clone() for t in first]                # geometry of static (read by the captured forward)
            torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                       # eager warm-up on a side stream (allocator, cuBLAS handles)
            for _ in range(warmup):
                self._step_body()
        torch.cuda.current_stream().wait_stream(side)
        dev0 = self.static[0].device
        torch.matmul(torch.zeros(1, 1, device=dev0), torch.zeros(1, 1, device=dev0))    # warmup = 0: the cuBLAS handle must exist before the capture
        torch.cuda.synchronize()
        with torch.cuda.graph(self.graph):
            self.loss = self._step_body()
        self._staged = None                                  # prefetch(): (host tensors, device staging, ready event)
        self._staging = None
        self._copy_stream = None
        self._consumed = None

    def prefetch(self, *host_batch):
        """Start the host -> device copy of the NEXT batch (pinned host tensors) on a copy stream, so that it overlaps the
        step that is running; the next __call__ with these very tensors takes the staged copy (one device-side copy into
        the static inputs) instead of copying from the host on the critical path -- what a prefetching DataLoader does for
        an eager loop.  Any other call falls back to the plain copy."""
        dev = self.static[0].device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._staging = [torch.empty_like(t) for t in self.static]
        cs = self._copy_stream
        if self._consumed is not None:
            cs.wait_event(self._consumed)                    # the staging buffers have been read by the previous hand-over
        with torch.cuda.stream(cs):
            for dst, src in zip(self._staging, host_batch):
                dst.copy_(src, non_blocking=True)
            ready = cs.record_event()
        self._staged = (tuple(host_batch), ready)

    def _take(self, dst_list, batch):
        """Copy `batch` into dst_list: from the staged device copy when it is the prefetched batch, else from where it lies."""
        if self._staged is not None and len(self._staged[0]) == len(batch) and all(a is b for a, b in zip(self._staged[0], batch)):
            torch.cuda.current_stream().wait_event(self._staged[1])
            torch._foreach_copy_(dst_list, self._staging)
            self._consumed = torch.cuda.current_stream().record_event()
            self._staged = None
            return
        for dst, src in zip(dst_list, batch):
            dst.copy_(src, non_blocking=True)

    def _step_body(self):
        from . import ops
        tmp = ev = None
        if self.geometry_fn is not None:
            torch._foreach_copy_(self.geo_cur, self.geo_next)       # the geometry of `static`, computed during the previous step
            aux = ops.aux_stream(self.static[0].device, which=1)
            tmp = self.geometry_fn(self.model, *self.static_next, stream=aux)      # next batch's geometry, side stream
            ev = aux.record_event()
        self.bucket.zero()
        if self.geometry_fn is not None:
            loss = self.forward_fn(self.model, *self.static, geometry=self.geo_cur)
        else:
            loss = self.forward_fn(self.model, *self.static)
        loss.backward()
        ops.join_aux()                                       # side-stream index work is consumed by the backward; belt and braces
        self.bucket.all_reduce_mean()
        self.opt.step()
        if tmp is not None:
            torch.cuda.current_stream().wait_event(ev)
            torch._foreach_copy_(self.geo_next, [t for t in tmp if isinstance(t, torch.Tensor)])
            del tmp
        return loss.detach()

    def __call__(self, *batch):
        if self.geometry_fn is not None:
            for cur, nxt in zip(self.static, self.static_next):
                cur.copy_(nxt, non_blocking=True)            # the batch handed in by the previous call becomes current
            self._take(self.static_next, batch)
        else:
            self._take(self.static, batch)
        self.graph.replay()
        return self.loss

    def flush(self):
        """Train on the batch still in the pipeline (geometry_fn mode): replays once more with the same next batch."""
        if self.geometry_fn is None:
            return None
        for cur, nxt in zip(self.static, self.static_next):
            cur.copy_(nxt, non_blocking=True)
        self.graph.replay()
        return self.loss
