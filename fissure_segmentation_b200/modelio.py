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

    def predict_full_pointcloud(self, pc, sample_points=1024, n_runs_min=50):
        """Ensemble prediction on a cloud larger than the training size: 4/5 of the runs draw random
        subsets, the remaining 1/5 target points no subset has touched yet (point_seg_net.py:21-48)."""
        n_total = pc.shape[-1]
        n_fill = n_runs_min // 5
        n_random = n_runs_min - n_fill
        acc = torch.zeros(pc.shape[0], self.num_classes, *pc.shape[2:], device=pc.device)
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
            for r in range(n_fill):
                part_unseen = unseen[order[r * n_unseen_per_run:(r + 1) * n_unseen_per_run]]
                part_seen = torch.randperm(len(seen), device=pc.device)[:n_seen_per_run]
                sub = torch.cat((part_unseen, part_seen), dim=0)
                acc[..., sub] += torch.softmax(self(pc[..., sub]).float(), dim=1)
            if (acc.sum(1) == 0).sum() != 0:
                warnings.warn('NOT ALL POINTS HAVE BEEN SEEN')

        return torch.softmax(acc, dim=1)
