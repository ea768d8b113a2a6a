"""Trainer-side glue around the hot path: student / teacher pair with a multi-tensor EMA update, the
phase-1 / phase-2 SSL losses, synthetic loaders for the three training modes, and a data-parallel training step.  Pure
host / library code: it carries no kernel of its own."""
from .momentum import MomentumModel
from .losses import ClusteredPrototypeLoss, ContrastivePairLoss
from .synthetic import synthetic_loader_multi_view, synthetic_loader_students_teacher, synthetic_loader_downstream

__all__ = ['MomentumModel', 'ClusteredPrototypeLoss', 'ContrastivePairLoss', 'synthetic_loader_multi_view',
           'synthetic_loader_students_teacher', 'synthetic_loader_downstream']
