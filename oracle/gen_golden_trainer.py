"""Golden vectors for the trainer-side glue from the LIVE reference (test infrastructure only):
    python -m oracle.gen_golden_trainer
ClusteredPrototypeLoss (losses/clustered_prototype_loss.py:13-206), ContrastivePairLoss (losses/contrastive_pair_loss.py)
and MomentumModel.update_teacher (momentum_model/momentum_model.py:27-36), float64, inputs from oracle.gen_golden_model.
det_tensor, torch seed fixed before every loss call (the prototype loss draws its crop jitter from the global generator)."""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402
from oracle.gen_golden_model import det_tensor, GOLDEN_DIR  # noqa: E402

PROTO_CASE = dict(B=2, C=6, teacher=(16, 16, 8), students=((16, 16, 8), (12, 12, 8)), rf=4.0, iters=3, fwhm=16.0, seed=77)


def proto_inputs():
    c = PROTO_CASE
    # float32: the reference builds its sampling grid in float32 (clustered_prototype_loss.py:162-165), float64 inputs fail
    grid = lambda size, off: torch.stack(torch.meshgrid(*[torch.arange(s, dtype=torch.float32) + o for s, o in zip(size, off)],
                                                        indexing='ij'), dim=0)
    emb_t = det_tensor((c["B"], c["C"], *c["teacher"]), 1.5).float()
    coord_t = grid(c["teacher"], (0, 0, 0)).unsqueeze(0).expand(c["B"], -1, -1, -1, -1).contiguous()
    emb_s, coord_s = [], []
    for i, size in enumerate(c["students"]):
        off = tuple((t - s) // 2 for t, s in zip(c["teacher"], size))
        emb_s.append(det_tensor((c["B"], c["C"], *size), 2.5 + i).float().requires_grad_(True))
        coord_s.append(grid(size, off).unsqueeze(0).expand(c["B"], -1, -1, -1, -1).contiguous())
    return emb_s, emb_t, coord_s, coord_t


def main():
    ref_loader.load()
    losses = importlib.import_module(f"{ref_loader._PARENT}.losses.clustered_prototype_loss")
    pair = importlib.import_module(f"{ref_loader._PARENT}.losses.contrastive_pair_loss")
    mm = importlib.import_module(f"{ref_loader._PARENT}.momentum_model.momentum_model")
    res = {}
    c = PROTO_CASE
    emb_s, emb_t, coord_s, coord_t = proto_inputs()
    torch.manual_seed(c["seed"])
    loss = losses.ClusteredPrototypeLoss(reduction_factor=c["rf"], k_means_iterations=c["iters"], fwhm=c["fwhm"])(
        emb_s, emb_t, coord_s, coord_t)
    loss.backward()
    res["proto.loss"] = loss.detach().numpy()
    for i, e in enumerate(emb_s):
        res[f"proto.grad{i}"] = e.grad.numpy()
    x_i, x_j = det_tensor((4, 16), 3.5).requires_grad_(True), det_tensor((4, 16), 4.5).requires_grad_(True)
    l2 = pair.ContrastivePairLoss(bs=4, temp=0.5).double()(x_i, x_j)
    l2.backward()
    res["pair.loss"], res["pair.grad_i"], res["pair.grad_j"] = l2.detach().numpy(), x_i.grad.numpy(), x_j.grad.numpy()

    class Net(torch.nn.Module):
        def __init__(self, conf):
            super().__init__()
            self.a = torch.nn.Linear(5, 3)
            self.b = torch.nn.LayerNorm(3)

    m = mm.MomentumModel(types.SimpleNamespace(tau=0.9), Net).double()
    with torch.no_grad():
        for k, (n, p) in enumerate(m.named_parameters()):
            p.copy_(det_tensor(p.shape, 10.0 + k))
    for _ in range(3):
        m.update_teacher()
    for n, p in m.net_teacher.named_parameters():
        res["ema." + n] = p.detach().numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "trainer_glue.npz"), **res)
    print("wrote trainer_glue", {k: v.shape for k, v in res.items()})


if __name__ == "__main__":
    main()
