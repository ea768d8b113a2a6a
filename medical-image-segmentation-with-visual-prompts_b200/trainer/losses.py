"""SSL losses of the reference trainers, restated for batched execution.

ClusteredPrototypeLoss (reference losses/clustered_prototype_loss.py:13-206, phase 2): teacher embeddings are
sub-sampled on a regular grid, soft k-means prototypes are refined with position-weighted assignments, and every
student view is trained with a cross entropy between its prototype similarities and the assignment of its spatially
closest teacher sample.  ContrastivePairLoss (losses/contrastive_pair_loss.py:6-31, phase 1): NT-Xent over two views.

Same constructor arguments, forward signatures, random-number consumption (one `torch.randint(0, ceil(rf), (6,))` per
student view, in order) and results as the reference (tests/test_trainer_cpu.py, goldens from the live reference).  What
differs is HOW: the reference loops over the batch with boolean-mask indexing (:80-89, a host synchronisation per
sample); here the masked means are weighted sums over the whole batch, so the loss is sync-free and graph-capturable.
"""
from __future__ import annotations

import math
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F


def _flat(t):                       # [B, C, H, W, D] -> [B, HWD, C]
    return t.flatten(2).transpose(1, 2)


def _regular_subsample(emb, coord, reduction_factor, jitter=False):
    """Bilinear samples of `emb` and `coord` on a regular grid of floor(size / rf) points per axis (identity affine grid,
    reflection padding, align_corners=False; reference :160-204).  jitter: crop a random margin of < ceil(rf) voxels first."""
    size = [max(int(s // reduction_factor), 1) for s in emb.shape[2:]]
    with torch.no_grad():
        theta = torch.eye(3, 4, device=emb.device).unsqueeze(0)
        grid = F.affine_grid(theta, [1, 1, *size], align_corners=False).expand(emb.shape[0], -1, -1, -1, -1)
        if jitter:
            j = torch.randint(low=0, high=int(math.ceil(reduction_factor)), size=(6,))
    if jitter:
        sl = (slice(None), slice(None), slice(int(j[0]), emb.shape[2] - int(j[1])), slice(int(j[2]), emb.shape[3] - int(j[3])),
              slice(int(j[4]), emb.shape[4] - int(j[5])))
        emb, coord = emb[sl], coord[sl]
    kw = dict(mode='bilinear', padding_mode='reflection', align_corners=False)
    return _flat(F.grid_sample(emb, grid, **kw)), F.grid_sample(coord, grid, **kw)


def _distances(coord_x, coord_y):
    """Euclidean distances between two coordinate grids [B, 3, ...] -> [B, Nx, Ny]."""
    x, y = _flat(coord_x), _flat(coord_y)
    return torch.linalg.norm(x[:, :, None, :] - y[:, None, :, :], ord=2, dim=-1)


def _position_weight(coord_x, coord_y, fwhm):
    sigma2 = (fwhm / 2.355) ** 2                     # FWHM ~= 2.355 sigma
    return torch.exp(-(_distances(coord_x, coord_y) ** 2 / (2 * sigma2)))


def _soft_assign(emb_x_n, emb_p_n, temp):
    return torch.softmax(torch.einsum('bnc,bpc->bnp', emb_x_n, emb_p_n) / temp, dim=-1)


def _cluster(emb_p, coord_p, emb_t, coord_t, n_iter, temp, fwhm):
    """Position-weighted soft k-means (reference :94-140): returns prototypes, their coordinates and the final weighted
    teacher -> prototype assignment."""
    emb_t_n = F.normalize(emb_t, p=2, dim=-1)
    emb_p_n = F.normalize(emb_p, p=2, dim=-1)
    grid_shape = coord_p.shape[2:]
    coord_t_flat = _flat(coord_t)
    for _ in range(n_iter):
        w = _soft_assign(emb_t_n, emb_p_n, temp) * _position_weight(coord_t, coord_p, fwhm)
        mass = w.sum(dim=1).unsqueeze(-1)
        emb_p = torch.einsum('bnp,bnc->bpc', w, emb_t) / mass
        emb_p_n = F.normalize(emb_p, p=2, dim=-1)
        coord_p = (torch.einsum('bnp,bnc->bpc', w, coord_t_flat) / mass).transpose(1, 2).unflatten(2, grid_shape)
    w = _soft_assign(emb_t_n, emb_p_n, temp) * _position_weight(coord_t, coord_p, fwhm)
    return emb_p, coord_p, w


def _assignment_loss(emb_z, coord_z, coord_t, emb_p, sim_t_p, temp, max_dist=4.0):
    """Per sample: mean over the student samples whose closest teacher sample lies within max_dist of
    -sum_p assignment(closest teacher sample)[p] * log softmax(student . prototypes / temp)[p]   (reference :64-91)."""
    sim = _soft_assign(F.normalize(emb_z, p=2, dim=-1), F.normalize(emb_p, p=2, dim=-1), temp)      # [B, N, P]
    dmin, closest = _distances(coord_z, coord_t).min(dim=-1)                                         # [B, N]
    target = torch.gather(sim_t_p, 1, closest.unsqueeze(-1).expand(-1, -1, sim_t_p.shape[-1]))       # [B, N, P]
    ce = -(target * torch.clamp(torch.log(sim + 1e-16), min=-1e3, max=-0.)).sum(dim=-1)              # [B, N]
    near = (dmin <= max_dist).to(ce.dtype)
    return (ce * near).sum(dim=1) / near.sum(dim=1)          # (0 / 0 = nan for a sample without a near pair, as the reference)


class ClusteredPrototypeLoss(nn.Module):
    def __init__(self, reduction_factor: float = 8.0, k_means_iterations: int = 3, fwhm: float = 128.0):
        super().__init__()
        self.reduction_factor = reduction_factor
        self.k_means_iterations = k_means_iterations
        self.fwhm = fwhm

    def forward(self, emb_s: List[torch.Tensor], emb_t: torch.Tensor, coord_s: List[torch.Tensor], coord_t: torch.Tensor,
                temp_s: float = 0.066, temp_t: float = 0.033):
        rf = self.reduction_factor
        emb_p, coord_p = _regular_subsample(emb_t, coord_t, rf * 2)
        emb_ts, coord_ts = _regular_subsample(emb_t, coord_t, rf)
        students = [_regular_subsample(e, c, rf, jitter=True) for e, c in zip(emb_s, coord_s)]
        emb_p, coord_p, sim_t_p = _cluster(emb_p, coord_p, emb_ts, coord_ts, self.k_means_iterations, temp_t, self.fwhm)
        total = torch.zeros((), device=emb_s[0].device)
        for emb_z, coord_z in students:
            total = total + _assignment_loss(emb_z, coord_z, coord_ts, emb_p, sim_t_p, temp_s).mean()
        return total


class ContrastivePairLoss(nn.Module):
    """NT-Xent over two batches of view embeddings [bs, C] (reference contrastive_pair_loss.py:6-31)."""

    def __init__(self, bs, temp=0.5):
        super().__init__()
        self.bs = bs
        self.register_buffer("temp", torch.tensor(temp))
        self.register_buffer("neg_mask", (~torch.eye(bs * 2, bs * 2, dtype=torch.bool)).float())

    def forward(self, x_i, x_j):
        z = F.normalize(torch.cat([F.normalize(x_i, dim=1), F.normalize(x_j, dim=1)]), dim=1)
        sim = z @ z.t()                                                   # cosine similarity of unit vectors
        pos = torch.cat([torch.diag(sim, self.bs), torch.diag(sim, -self.bs)], dim=0) / self.temp
        log_neg = torch.log((self.neg_mask.to(z.device) * torch.exp(sim / self.temp)).sum(dim=1))
        return (log_neg - pos).sum() / (2 * self.bs)
