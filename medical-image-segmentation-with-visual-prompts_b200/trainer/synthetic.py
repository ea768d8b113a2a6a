"""Synthetic stand-ins for the reference's MONAI data loaders (datasets/utils.py:9,53,100): iterables of batches with the
dict keys and shapes the three trainers read, filled with CT-like values in [0, 1) (the reference scales HU to [0, 1],
datasets/transforms.py:142-147).  Batches are created in PINNED host memory, so that a training loop can overlap the
host -> device copy with the previous step (graphs.InputPrefetcher).  Deterministic per (seed, rank)."""
from __future__ import annotations

from typing import Iterator, Sequence

import torch


def _coord_grid(size: Sequence[int]) -> torch.Tensor:
    """Voxel-index coordinate grid [3, H, W, D] (reference datasets/transforms.py:337-344)."""
    axes = [torch.arange(s, dtype=torch.float32) for s in size]
    return torch.stack(torch.meshgrid(*axes, indexing='ij'), dim=0)


def _pin(t: torch.Tensor) -> torch.Tensor:
    return t.pin_memory() if torch.cuda.is_available() else t


def synthetic_loader_multi_view(steps: int, batch: int, patch: Sequence[int] = (96, 96, 96), channels: int = 1, seed: int = 1234,
                                rank: int = 0) -> Iterator[dict]:
    """Phase 1 (multi_view.py:118-136): {'image': [B, C, *patch]}; the trainer derives its two views from it."""
    gen = torch.Generator().manual_seed(seed + rank)
    for _ in range(steps):
        yield {'image': _pin(torch.rand((batch, channels, *patch), generator=gen))}


def synthetic_loader_students_teacher(steps: int, batch: int, teacher_size: Sequence[int] = (96, 96, 96),
                                      student_sizes: Sequence[Sequence[int]] = ((96, 96, 96), (96, 96, 96)), channels: int = 1,
                                      seed: int = 1234, rank: int = 0) -> Iterator[dict]:
    """Phase 2 (students_teacher.py:152-161): a teacher crop, one crop per student size, and the voxel coordinates of
    every crop in the frame of the teacher crop (students are random sub-crops of it)."""
    gen = torch.Generator().manual_seed(seed + rank)
    full = _coord_grid(teacher_size)
    for _ in range(steps):
        x_t = torch.rand((batch, channels, *teacher_size), generator=gen)
        out = {'image_teacher': _pin(x_t), 'coord_teacher': _pin(full.unsqueeze(0).expand(batch, -1, -1, -1, -1).contiguous()),
               'image_students': [], 'coord_students': []}
        for size in student_sizes:
            off = [int(torch.randint(0, t - s + 1, (1,), generator=gen)) for t, s in zip(teacher_size, size)]
            sl = tuple(slice(o, o + s) for o, s in zip(off, size))
            out['image_students'].append(_pin(x_t[(slice(None), slice(None), *sl)].contiguous()))
            out['coord_students'].append(_pin(full[(slice(None), *sl)].unsqueeze(0).expand(batch, -1, -1, -1, -1).contiguous()))
        yield out


def synthetic_loader_downstream(steps: int, batch: int, patch: Sequence[int] = (96, 96, 96), channels: int = 1, classes: int = 2,
                                seed: int = 1234, rank: int = 0) -> Iterator[dict]:
    """Downstream few-shot (segmentation.py:98-103): {'image': [B, C, *patch], 'mask': [B, 1, *patch] integer labels}."""
    gen = torch.Generator().manual_seed(seed + rank)
    for _ in range(steps):
        yield {'image': _pin(torch.rand((batch, channels, *patch), generator=gen)),
               'mask': _pin(torch.randint(0, classes, (batch, 1, *patch), generator=gen))}
