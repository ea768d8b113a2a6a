"""Student / teacher pair for phase 2 (reference momentum_model/momentum_model.py:4-36).

Same constructor (`conf.tau`, an `architecture` class taking `conf=`), same `copy_state_dict`, `forward(x_students,
x_teacher)` and `update_teacher` semantics: theta_t <- tau * theta_t + (1 - tau) * theta_s over `named_parameters()` in
order.  The reference walks ~600 parameter pairs in a Python loop and REBINDS `param_teacher.data` to a fresh tensor each
step (:27-36): ~1800 tiny launches and allocations per step at 9.5 M parameters.  Here the update is TWO multi-tensor
launches in place (`torch._foreach_mul_`, `torch._foreach_add_`), which also keeps the teacher's storage stable -- a
CUDA-graph-captured teacher forward stays valid across updates.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class MomentumModel(nn.Module):
    def __init__(self, conf, architecture):
        super().__init__()
        self.tau = conf.tau
        self.net_student = architecture(conf=conf)
        self.net_teacher = architecture(conf=conf)
        self._pairs = None

    def _param_pairs(self):
        if self._pairs is None:
            s = [p for _, p in self.net_student.named_parameters()]
            t = [p for _, p in self.net_teacher.named_parameters()]
            if len(s) != len(t):
                raise RuntimeError("MomentumModel: student and teacher have different parameter lists")
            self._pairs = (s, t)
        return self._pairs

    def copy_state_dict(self):
        s, t = self._param_pairs()
        with torch.no_grad():
            torch._foreach_copy_([p.data for p in t], [p.data for p in s])
        for p in t:
            p.requires_grad = False

    def forward(self, x_students, x_teacher):
        return [self.net_student(x) for x in x_students], self.net_teacher(x_teacher)

    @torch.no_grad()
    def update_teacher(self):
        s, t = self._param_pairs()
        td = [p.data for p in t]
        torch._foreach_mul_(td, self.tau)
        torch._foreach_add_(td, [p.data for p in s], alpha=1.0 - self.tau)
