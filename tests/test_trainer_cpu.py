"""Trainer-side glue against the live reference's float64 results (tests/golden/trainer_glue.npz, made by
oracle/gen_golden_trainer.py): the batched prototype loss, NT-Xent, the multi-tensor EMA, the synthetic loaders.  Pure torch
host code: runs on CPU."""
import types

import numpy as np
import torch

import pwa_b200
from pwa_b200 import trainer as T
from oracle import gen_golden_trainer as G
from oracle.gen_golden_model import det_tensor
from tests.util import load_npz, rel_linf


def test_clustered_prototype_loss_vs_reference():
    d = load_npz("trainer_glue")
    c = G.PROTO_CASE
    emb_s, emb_t, coord_s, coord_t = G.proto_inputs()
    torch.manual_seed(c["seed"])                       # same crop jitter as the reference drew
    loss = T.ClusteredPrototypeLoss(reduction_factor=c["rf"], k_means_iterations=c["iters"], fwhm=c["fwhm"])(
        emb_s, emb_t, coord_s, coord_t)
    loss.backward()
    # float32 on both sides (the reference's sampling grid is float32); batched vs per-sample summation order
    assert abs(loss.item() - float(d["proto.loss"])) < 1e-5 * max(1.0, abs(float(d["proto.loss"])))
    for i, e in enumerate(emb_s):
        assert rel_linf(e.grad, torch.from_numpy(d[f"proto.grad{i}"])) < 1e-4


def test_contrastive_pair_loss_vs_reference():
    d = load_npz("trainer_glue")
    x_i, x_j = det_tensor((4, 16), 3.5).requires_grad_(True), det_tensor((4, 16), 4.5).requires_grad_(True)
    loss = T.ContrastivePairLoss(bs=4, temp=0.5).double()(x_i, x_j)
    loss.backward()
    assert abs(loss.item() - float(d["pair.loss"])) < 1e-12
    assert rel_linf(x_i.grad, torch.from_numpy(d["pair.grad_i"])) < 1e-10
    assert rel_linf(x_j.grad, torch.from_numpy(d["pair.grad_j"])) < 1e-10


def test_momentum_model_multi_tensor_ema_vs_reference():
    d = load_npz("trainer_glue")

    class Net(torch.nn.Module):
        def __init__(self, conf):
            super().__init__()
            self.a = torch.nn.Linear(5, 3)
            self.b = torch.nn.LayerNorm(3)

    m = T.MomentumModel(types.SimpleNamespace(tau=0.9), Net).double()
    with torch.no_grad():
        for k, (n, p) in enumerate(m.named_parameters()):
            p.copy_(det_tensor(p.shape, 10.0 + k))
    ptrs = [p.data_ptr() for p in m.net_teacher.parameters()]
    for _ in range(3):
        m.update_teacher()
    for n, p in m.net_teacher.named_parameters():
        assert rel_linf(p, torch.from_numpy(d["ema." + n])) < 1e-14, n
    assert ptrs == [p.data_ptr() for p in m.net_teacher.parameters()]      # in place: storage stays put (graph-safe)
    m.copy_state_dict()
    for ps, pt in zip(m.net_student.parameters(), m.net_teacher.parameters()):
        assert torch.equal(ps, pt) and not pt.requires_grad


def test_synthetic_loaders_shapes_and_determinism():
    a = list(T.synthetic_loader_multi_view(2, 3, patch=(8, 8, 8)))
    b = list(T.synthetic_loader_multi_view(2, 3, patch=(8, 8, 8)))
    assert a[0]['image'].shape == (3, 1, 8, 8, 8) and torch.equal(a[1]['image'], b[1]['image'])
    assert not torch.equal(a[0]['image'], list(T.synthetic_loader_multi_view(1, 3, patch=(8, 8, 8), rank=1))[0]['image'])
    st = next(iter(T.synthetic_loader_students_teacher(1, 2, teacher_size=(8, 8, 4), student_sizes=((8, 8, 4), (4, 4, 4)))))
    assert st['image_teacher'].shape == (2, 1, 8, 8, 4) and st['coord_teacher'].shape == (2, 3, 8, 8, 4)
    assert st['image_students'][1].shape == (2, 1, 4, 4, 4) and st['coord_students'][1].shape == (2, 3, 4, 4, 4)
    # a student crop is the teacher crop at the coordinates it reports
    c = st['coord_students'][1][0].long()
    assert torch.equal(st['image_students'][1][0, 0], st['image_teacher'][0, 0][c[0], c[1], c[2]])
    ds = next(iter(T.synthetic_loader_downstream(1, 2, patch=(8, 8, 8), classes=3)))
    assert ds['mask'].shape == (2, 1, 8, 8, 8) and int(ds['mask'].max()) <= 2
