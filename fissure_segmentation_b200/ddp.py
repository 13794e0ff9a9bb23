"""Data-parallel harness: one process per GPU, batch-sharded clouds, gradient all-reduce overlapped
with backward. The reference has no distributed code; this sits beside its trainer.

Parameters and gradients live in two flat fp32 buffers (each nn.Parameter is a view). Gradients are
grouped into a few buckets in reverse registration order (the order backward produces them); when the
last gradient of a bucket has been accumulated, an asynchronous all-reduce (NCCL over NVLink on GPUs,
gloo in the CPU tests) is issued on that bucket while the rest of backward is still running. BatchNorm
statistics stay per rank (standard DDP semantics).
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib


class FlatDataParallel:
    def __init__(self, module, n_buckets=2, process_group=None, broadcast=True, tail_share=None, fused_tail=False):
        self.module = module
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        params = [p for p in module.parameters() if p.requires_grad]
        if not params:
            raise ValueError("module has no trainable parameters")
        dev, dt = params[0].device, params[0].dtype
        # reverse registration order ~ order in which backward finishes the gradients
        order = list(reversed(params))
        total = sum(p.numel() for p in order)
        self.flat_param = torch.empty(total, dtype=dt, device=dev)
        # fused_tail: the gradient buffer lives in symmetric memory (every rank maps every rank's buffer over NVLink), so
        # the LAST bucket needs no collective call: FlatAdam reads the peers' slices directly and reduces + updates in one
        # kernel (fs_adam_step_peers). The gradients of that bucket's parameters then stay rank-local in `.grad`.
        # fused_tail="all" makes the whole parameter set that bucket (2.5 MB for DGCNNSeg: a one-shot read of all peers).
        self.symm = None
        self.flat_grad = None
        if fused_tail == "all":       # every gradient through peer memory: one bucket, no collective call in the step
            n_buckets = 1
        if fused_tail and self.world > 1 and dev.type == "cuda" and dt == torch.float32:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                buf = symm_mem.empty(total, dtype=dt, device=dev)
                self.symm = symm_mem.rendezvous(buf, process_group if process_group is not None else dist.group.WORLD)
                buf.zero_()
                self.flat_grad = buf
            except Exception as exc:      # no peer access / symmetric memory on this system: the NCCL path below
                import warnings
                warnings.warn("FlatDataParallel: symmetric memory unavailable (%s); tail bucket uses all_reduce" % (exc,))
                self.symm = None
        if self.flat_grad is None:
            self.flat_grad = torch.zeros(total, dtype=dt, device=dev)
        self._slices = []
        self._grad_views = []
        off = 0
        with torch.no_grad():
            for p in order:
                n = p.numel()
                self.flat_param[off:off + n].copy_(p.detach().reshape(-1))
                p.data = self.flat_param[off:off + n].view_as(p)
                self._grad_views.append(self.flat_grad[off:off + n].view_as(p))
                p.grad = None
                self._slices.append((off, n))
                off += n
        self.params = order
        # contiguous buckets of roughly equal size; with tail_share the LAST bucket (the gradients backward produces
        # last - its all-reduce is the only one nothing can overlap) holds about that fraction of the parameters and
        # the others share the rest: a small tail keeps the exposed collective latency-bound instead of size-bound
        n_buckets = max(1, min(n_buckets, len(order)))
        if tail_share is not None and n_buckets > 1:
            head = total * (1.0 - float(tail_share)) / (n_buckets - 1)
            targets = [head] * (n_buckets - 1) + [total]
        else:
            targets = [total / n_buckets] * n_buckets
        self.buckets = []          # (start, end, n_params)
        self._bucket_members = []  # parameter positions of every bucket
        self._bucket_of = {}
        start, count, b, first = 0, 0, 0, 0
        for i, (o, n) in enumerate(self._slices):
            self._bucket_of[id(order[i])] = b
            count += 1
            end = o + n
            target = targets[b]
            if (end - start >= target and b < n_buckets - 1) or i == len(order) - 1:
                self.buckets.append((start, end, count))
                self._bucket_members.append(list(range(first, i + 1)))
                start, count, b, first = end, 0, b + 1, i + 1
        self._ready = [0] * len(self.buckets)
        self._sent = [False] * len(self.buckets)
        self._works = []
        self._reduced = False      # True between finish_backward() and zero_grad(): .grad holds all-reduced sums
        self.skip_collectives = False   # measurement only: leaves the gradients un-reduced (what do the collectives cost?)
        if self.world > 1 and broadcast:
            dist.broadcast(self.flat_param, src=0, group=self.group)
        # Gradients are NOT accumulated into the flat buffer by autograd (that costs one add kernel per parameter
        # and a memset per step): zero_grad() drops them, autograd assigns the fresh tensors, and when the last
        # gradient of a bucket has arrived ONE multi-tensor copy moves the bucket into the flat buffer, followed by
        # its asynchronous all-reduce.
        for p in order:
            p.register_post_accumulate_grad_hook(self._hook)

    def _flush(self, b):
        s, e, _ = self.buckets[b]
        src, dst, missing = [], [], False
        for i in self._bucket_members[b]:
            g = self.params[i].grad
            if g is None:
                missing = True
            elif g.data_ptr() != self._grad_views[i].data_ptr():
                src.append(g)
                dst.append(self._grad_views[i])
        if missing:                 # parameters without a gradient this step count as zero
            self.flat_grad[s:e].zero_()
        if src:
            fast = all(g.is_cuda and g.dtype == torch.float32 and g.is_contiguous() for g in src) and \
                self.flat_grad.dtype == torch.float32
            if fast:
                # one launch for the whole bucket (the multi-tensor ATen copy takes ~17 us per bucket of ~15 tensors)
                n = len(src)
                sp = (ctypes.c_void_p * n)(*[g.data_ptr() for g in src])
                dp = (ctypes.c_void_p * n)(*[d.data_ptr() for d in dst])
                cn = (ctypes.c_longlong * n)(*[g.numel() for g in src])
                _lib.call("fs_multi_copy_f32", self.flat_grad, n, sp, dp, cn)
            else:
                torch._foreach_copy_(dst, src)
        for i in self._bucket_members[b]:
            if self.params[i].grad is not None:
                self.params[i].grad = self._grad_views[i]          # .grad shows the (to be) reduced gradient
        self._sent[b] = True
        if self.world > 1 and not (self.symm is not None and b == len(self.buckets) - 1) and not self.skip_collectives:
            self._works.append(dist.all_reduce(self.flat_grad[s:e], op=dist.ReduceOp.SUM, group=self.group,
                                               async_op=True))

    def _hook(self, p):
        if self._reduced:
            # Contract: one backward per zero_grad(). After finish_backward() every .grad is a view of the flat
            # buffer that already holds the SUM over ranks; a second backward would accumulate local gradients on
            # top of it and reduce the mixture again (silently wrong: ~ world x old + new).
            raise RuntimeError("FlatDataParallel: backward() after finish_backward() without zero_grad(); gradient "
                               "accumulation over several backward passes is not supported - call dp.zero_grad() "
                               "at the start of every step")
        b = self._bucket_of[id(p)]
        self._ready[b] += 1
        if self._ready[b] == self.buckets[b][2] and not self._sent[b]:
            self._flush(b)

    def __call__(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def zero_grad(self):
        for p in self.params:
            p.grad = None
        self._reduced = False

    def finish_backward(self):
        """Flush the buckets whose parameters did not all receive a gradient, wait for the outstanding bucket
        all-reduces; flat_grad then holds the SUM over ranks (the 1/world average is folded into the optimiser's
        grad_scale)."""
        for b in range(len(self.buckets)):
            if not self._sent[b]:
                self._flush(b)
        for w in self._works:
            w.wait()
        self._works = []
        self._ready = [0] * len(self.buckets)
        self._sent = [False] * len(self.buckets)
        self._reduced = True


class FlatAdam:
    """torch.optim.Adam(lr, weight_decay) semantics (model_trainer.py:57) as one fused kernel over the
    flat buffers of FlatDataParallel. Step count and learning rate live on the device so the whole
    training step can be captured in a CUDA graph."""

    def __init__(self, dp, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5):
        self.dp = dp
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.exp_avg = torch.zeros_like(dp.flat_param)
        self.exp_avg_sq = torch.zeros_like(dp.flat_param)
        self.dyn = torch.tensor([0.0, lr], dtype=torch.float32, device=dp.flat_param.device)   # [step, lr]
        self._inc = torch.tensor([1.0, 0.0], dtype=torch.float32, device=dp.flat_param.device)
        self.tail_sum = None       # optional fp32 buffer (length of the tail bucket) that receives the peer-reduced gradient

    def set_lr(self, lr):
        self.lr = lr
        self.dyn[1] = lr

    def step(self):
        self.dyn.add_(self._inc)
        dp = self.dp
        hyper = (float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.weight_decay))
        if dp.symm is None:
            _lib.call("fs_adam_step", dp.flat_param, dp.flat_param, dp.flat_grad, self.exp_avg, self.exp_avg_sq,
                      dp.flat_param.numel(), *hyper, 0, 1.0 / dp.world, self.dyn)
            return
        # head buckets: reduced by NCCL during backward; tail bucket: reduced from peer memory inside the update kernel
        s, e, _ = dp.buckets[-1]
        if s > 0:
            _lib.call("fs_adam_step", dp.flat_param, dp.flat_param, dp.flat_grad, self.exp_avg, self.exp_avg_sq, s,
                      *hyper, 0, 1.0 / dp.world, self.dyn)
        dp.symm.barrier(channel=0)          # every rank's tail gradients are in its buffer
        _lib.call("fs_adam_step_peers", dp.flat_param, dp.flat_param[s:e], dp.symm.buffer_ptrs_dev, dp.world, s,
                  self.exp_avg[s:e], self.exp_avg_sq[s:e], e - s, *hyper, 0, 1.0 / dp.world, self.dyn, self.tail_sum)
        dp.symm.barrier(channel=1)          # no rank overwrites its gradients before every rank has read them
