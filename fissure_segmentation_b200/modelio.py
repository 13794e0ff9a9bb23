"""Config capture and checkpoint format of the reference's models (models/modelio.py:20-89,
models/point_seg_net.py:9-48), so reference checkpoints load into the B200 modules and
`type(model)(**model.config)` (train.py:505) rebuilds them.
"""
import functools
import inspect
import warnings
from abc import ABC, abstractmethod

import torch
from torch import nn


def store_config_args(init):
    """Decorator for __init__: records defaults, positional and keyword arguments in `self.config`
    (same dictionary the reference builds at models/modelio.py:20-48; uses getfullargspec because
    inspect.getargspec no longer exists)."""
    spec = inspect.getfullargspec(init)
    names, defaults = spec.args, spec.defaults or ()

    @functools.wraps(init)
    def wrapped(self, *args, **kwargs):
        cfg = dict(zip(names[len(names) - len(defaults):], defaults))
        cfg.update(zip(names[1:], args))
        cfg.update(kwargs)
        self.config = cfg
        return init(self, *args, **kwargs)

    return wrapped


class LoadableModel(nn.Module):
    """nn.Module whose constructor arguments travel with the weights: `save` writes
    {'config', 'model_state'} and `load` rebuilds cls(**config) (models/modelio.py:51-89)."""

    def __init__(self, *args, **kwargs):
        if not hasattr(self, "config"):
            raise RuntimeError("models that inherit from LoadableModel must decorate the constructor with "
                               "@store_config_args")
        super().__init__(*args, **kwargs)

    def save(self, path):
        state = {k: v for k, v in self.state_dict().items() if not k.endswith(".grid")}
        torch.save({"config": self.config, "model_state": state}, path)

    @classmethod
    def load(cls, path, device):
        ckpt = torch.load(path, map_location=torch.device(device))
        model = cls(**ckpt["config"])
        model.load_state_dict(ckpt["model_state"], strict=False)
        return model


class PointSegmentationModelBase(LoadableModel, ABC):
    """Base of the point segmentation nets (models/point_seg_net.py:9-48)."""

    @store_config_args
    def __init__(self, in_features, num_classes, **kwargs):
        super().__init__()
        self.in_features = in_features
        self.num_classes = num_classes

    @abstractmethod
    def forward(self, x):
        pass

    # ---- ensemble inference -------------------------------------------------------------------------------------
    inference_cuda_graph = True        # capture the batched eval forward once per (runs, channels, subset size)

    def _forward_subsets(self, pc, sub):
        """Logits (R, classes, S) of the R subsets `sub` (R, S) of the single cloud pc (1, C, n): ONE eval forward of
        batch R instead of R launch-bound B = 1 forwards. Eval-mode BatchNorm uses the running statistics, so the
        clouds of a batch do not interact and the result equals the sequential loop's."""
        x = pc[0][:, sub].permute(1, 0, 2).contiguous()                       # (R, C, S)
        if not (self.inference_cuda_graph and x.is_cuda and not torch.is_grad_enabled()):
            return self(x).float()
        # CUDA graph of the forward, keyed by the shape and by the storage of every parameter / buffer (a reloaded or
        # moved model re-captures; in-place weight updates are picked up by the replay)
        key = (tuple(x.shape), x.dtype, x.device, tuple(t.data_ptr() for t in self.state_dict().values()))
        cache = self.__dict__.setdefault("_infer_graphs", {})
        entry = cache.get(key)
        if entry is None:
            if len(cache) >= 4 or any(k[3] != key[3] for k in cache):     # other weights storage: drop stale graphs
                cache.clear()
            static_x = x.clone()
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(torch.cuda.current_stream(x.device))
            with torch.cuda.stream(side):
                for _ in range(2):                                          # warm-up: lazy initialisations, arena sizing
                    self(static_x)
            torch.cuda.current_stream(x.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self(static_x).float()
            entry = cache[key] = (graph, static_x, static_out)
        graph, static_x, static_out = entry
        static_x.copy_(x)
        graph.replay()
        return static_out.clone()

    def _accumulate_runs(self, acc, pc, sub, dedupe=False):
        """acc[0, :, sub[r]] += softmax(forward(pc[..., sub[r]])) for every run r (point_seg_net.py:27-29, :43).
        dedupe: a fill run may hold the same point twice (`% len(left_out_pts)`, :38); the reference's `+=` is a
        non-accumulating index_put that keeps ONE of the duplicates' values, so only the first occurrence counts."""
        from . import _lib
        logits = self._forward_subsets(pc, sub).contiguous()
        R, classes, S = logits.shape
        if dedupe:
            srt, pos = sub.sort(dim=1, stable=True)
            dup = torch.zeros_like(sub, dtype=torch.bool)
            dup[:, 1:] = srt[:, 1:] == srt[:, :-1]
            sub = sub.masked_fill(torch.zeros_like(dup).scatter_(1, pos, dup), -1)      # the kernel skips index < 0
        _lib.call("fs_softmax_scatter_add", logits, logits, sub.contiguous(), R, classes, S, acc.shape[-1], acc[0])

    def predict_full_pointcloud(self, pc, sample_points=1024, n_runs_min=50):
        """Ensemble prediction on a cloud larger than the training size: 4/5 of the runs draw random
        subsets, the remaining 1/5 target points no subset has touched yet (point_seg_net.py:21-48).

        For a single CUDA cloud in eval mode the runs of each phase are batched: the random draws are made in the
        reference's order (so a seeded generator gives the sequential loop's subsets), then ONE forward of batch
        n_random (and one of batch n_fill) replaces the 50 B = 1 forwards, and one scatter kernel accumulates the
        probabilities. Any other case (training-mode BatchNorm couples the clouds of a batch) runs the loop."""
        n_total = pc.shape[-1]
        n_fill = n_runs_min // 5
        n_random = n_runs_min - n_fill
        batched = (pc.is_cuda and pc.shape[0] == 1 and pc.dim() == 3 and not self.training and n_random > 0
                   and self.num_classes <= 32 and sample_points <= n_total)
        acc = torch.zeros(pc.shape[0], self.num_classes, *pc.shape[2:], device=pc.device)
        if batched:
            sub = torch.stack([torch.randperm(n_total, device=pc.device)[:sample_points] for _ in range(n_random)])
            self._accumulate_runs(acc, pc, sub)
        else:
            for _ in range(n_random):
                sub = torch.randperm(n_total, device=pc.device)[:sample_points]
                acc[..., sub] += torch.softmax(self(pc[..., sub]).float(), dim=1)

        unseen = torch.nonzero(acc.sum(1) == 0)[..., 1]
        print(f'After {n_random} runs, {unseen.shape[0]} points have not been seen yet.')
        if unseen.shape[0] > 0:
            seen = torch.nonzero(acc.sum(1))[..., 1]
            n_unseen_per_run = sample_points // 2
            n_seen_per_run = sample_points - n_unseen_per_run
            order = torch.randperm(n_fill * n_unseen_per_run, device=pc.device) % len(unseen)
            subs = []
            for r in range(n_fill):
                part_unseen = unseen[order[r * n_unseen_per_run:(r + 1) * n_unseen_per_run]]
                part_seen = torch.randperm(len(seen), device=pc.device)[:n_seen_per_run]
                sub = torch.cat((part_unseen, part_seen), dim=0)
                if batched and sub.shape[0] == sample_points:
                    subs.append(sub)
                else:
                    acc[..., sub] += torch.softmax(self(pc[..., sub]).float(), dim=1)
            if subs:
                self._accumulate_runs(acc, pc, torch.stack(subs), dedupe=True)
            if (acc.sum(1) == 0).sum() != 0:
                warnings.warn('NOT ALL POINTS HAVE BEEN SEEN')

        return torch.softmax(acc, dim=1)
