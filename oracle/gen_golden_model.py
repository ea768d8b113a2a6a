"""Model-level golden vectors: the LIVE reference `SwinUnetR` (swin_unetr/swin_unetr.py) executed here on CPU.

Test infrastructure only.  Run in the build container (where /root/reference is mounted):
    python -m oracle.gen_golden_model
`swin_unetr.py:1` and `unet_blocks.py:2-3` import five MONAI symbols and MONAI is not installed (and not pinned by the
reference: no requirements file).  With the reference's default wiring (`unetr_res_block: none`, `unetr_up_block: swin`)
only three of them execute, with fixed arguments (unet_blocks.py:36-56): `get_act_layer("leakyrelu")`,
`get_norm_layer("batch", 3, ch)` and `Convolution(3, cin, cout, strides=1, kernel_size=3, conv_only=True)`.  The stub
below provides exactly those as `nn.LeakyReLU()`, `nn.BatchNorm3d(ch)` and a Sequential with one child `conv =
Conv3d(k=3, padding=1)` -- what MONAI builds for these arguments; parity of these three layers is therefore pinned to
torch's own modules, not to a MONAI release.  `UnetrBasicBlock` / `UnetrUpBlock` are never constructed.

Cases (small enough for CPU float64, all with prompting and training-mode BatchNorm statistics):
  cfg1        BASELINE.json configs[0]: feature_size 12, 64^3, batch 1, self_supervised_learning_encoder
  cfg3_small  configs[2] reduced: feature_size 12, 32x32x32 input, batch 2, self_supervised_learning_all
              (encoder + decoder prompting, 12 prompted blocks)
  cfg4_small  configs[3] reduced: downstream (frozen backbone), same size
Inputs, upstream gradients and EVERY parameter are deterministic closed-form functions of (tensor name, shape)
(`det_tensor` / `fill_params_deterministic` below, shared with tests/test_gpu_model.py), so the fixtures only hold the
reference's results: every output tensor (large volumes sub-sampled with a fixed stride) and the float64 gradients of a
fixed set of parameters incl. all prompt tokens.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def det_tensor(shape, salt):
    """Deterministic pseudo-random float64 tensor in [-1, 1): frac(sin(i * 12.9898 + salt) * 43758.5453) * 2 - 1."""
    n = int(np.prod(shape)) if len(shape) else 1
    v = np.sin(np.arange(n, dtype=np.float64) * 12.9898 + float(salt)) * 43758.5453
    return torch.from_numpy((v - np.floor(v)) * 2.0 - 1.0).reshape(tuple(shape))


def _salt(name):
    h = 0
    for ch in name:
        h = (h * 131 + ord(ch)) % 1000003
    return h * 0.001


def fill_params_deterministic(model):
    """Overwrite every parameter by a function of its NAME and shape (same on the reference and on our mirror, whatever
    their construction order): matrices / conv kernels uniform with the xavier bound of their shape, norm weights
    1 + 0.2 u, every other vector 0.1 u.  Buffers (BatchNorm running statistics, index tables) keep their defaults."""
    with torch.no_grad():
        for n, p in model.named_parameters():
            u = det_tensor(p.shape, _salt(n)).to(p.dtype)
            if p.dim() >= 2:
                fan_out, fan_in = p.shape[0], int(np.prod(p.shape[1:]))
                p.copy_(u * (6.0 / (fan_in + fan_out)) ** 0.5 * (3.0 if ".pe." in n else 1.0))
            elif "norm" in n and n.endswith("weight") or (n.endswith(".1.weight") or n.endswith(".0.weight")) and p.dim() == 1:
                p.copy_(1.0 + 0.2 * u)
            else:
                p.copy_(0.1 * u)


def _install_monai_stub():
    if "monai" in sys.modules:
        return

    def get_act_layer(name):
        assert name == "leakyrelu", name
        return nn.LeakyReLU()

    def get_norm_layer(name, spatial_dims, channels):
        assert name == "batch" and spatial_dims == 3
        return nn.BatchNorm3d(channels)

    class Convolution(nn.Sequential):
        def __init__(self, spatial_dims, in_channels, out_channels, strides=1, kernel_size=3, act=None, norm=None,
                     conv_only=False, is_transposed=False):
            assert spatial_dims == 3 and conv_only and not is_transposed
            ks = tuple(kernel_size) if not isinstance(kernel_size, int) else (kernel_size,) * 3
            super().__init__(OrderedDict(conv=nn.Conv3d(in_channels, out_channels, kernel_size=ks, stride=strides,
                                                        padding=tuple(k // 2 for k in ks))))

    class _NotBuilt(nn.Module):
        def __init__(self, *a, **k):
            raise RuntimeError("MONAI Unetr blocks are not part of the stub (non-default wiring)")

    mods = {n: types.ModuleType(n) for n in ("monai", "monai.networks", "monai.networks.blocks", "monai.networks.layers",
                                              "monai.networks.layers.utils")}
    mods["monai.networks.blocks"].UnetrBasicBlock = _NotBuilt
    mods["monai.networks.blocks"].UnetrUpBlock = _NotBuilt
    mods["monai.networks.blocks"].Convolution = Convolution
    mods["monai.networks.layers.utils"].get_act_layer = get_act_layer
    mods["monai.networks.layers.utils"].get_norm_layer = get_norm_layer
    sys.modules.update(mods)


def load_reference_model_class():
    _install_monai_stub()
    ref_loader.load()
    return importlib.import_module(f"{ref_loader._PARENT}.swin_unetr.swin_unetr").SwinUnetR


def model_conf(mode, feature=12, enc_prompt=True, dec_prompt=False):
    return types.SimpleNamespace(
        training_mode=mode, input_channels=1, depth_unet=3, hidden_channels=[feature, 2 * feature, 4 * feature, 8 * feature],
        input_patch_size=[2, 2, 2], unetr_res_block='none', unetr_up_block='swin', basic_block_res=True,
        num_heads_encoder=4, num_heads_decoder=4, attn_window_size=[8, 8, 4], pos_bias_embed_dim=64, use_checkpoint=False,
        attn_drop=0.0, proj_drop=0.0, max_prompts=1, tokens_per_prompt_encoder=64, tokens_per_prompt_decoder=64,
        use_encoder_prompting=enc_prompt, use_decoder_prompting=dec_prompt, contrastive_coding_dim=32,
        use_reconstruction=True, use_rotation_prediction=True, use_contrastive_learning=True, use_mutual_learning=False,
        output_channels_pretrain=5, output_channels_downstream=2)


CASES = [  # name, mode, input dims, batch, decoder prompting
    ("model_cfg1", 'self_supervised_learning_encoder', (64, 64, 64), 1, False),
    ("model_cfg3_small", 'self_supervised_learning_all', (32, 32, 32), 2, True),
    ("model_cfg4_small", 'downstream', (32, 32, 32), 2, True),
]

SUB = 4      # stride of the stored sub-sample of volumes with more than 2^15 elements


def subsample(t):
    return t[..., ::SUB, ::SUB, ::SUB] if t.dim() == 5 and t.numel() > (1 << 15) else t


def output_items(out):
    """(name, tensor) pairs of a model output dict in a fixed order; `out_vit` is a list."""
    items = []
    for k in sorted(out):
        if k == 'out_vit':
            items += [(f"out_vit.{i}", t) for i, t in enumerate(out[k][:-1])]      # the last entry is the input itself
        else:
            items.append((k, out[k]))
    return items


def grad_param_names(model):
    """Parameters whose gradients are stored (at most 20 000 elements each): every prompt token set, the patch embedding,
    per stage the first block's projection / norm / bias parameters, and the merge / conv / head layers around them."""
    keep = []
    for n, p in model.named_parameters():
        if p.numel() > 20000:
            continue
        if n.startswith("prompt_tokens.") or n.startswith("input_layer."):
            keep.append(n)
        elif ".swin_blocks.0." in n and any(s in n for s in ("to_q.weight", "proj.bias", "attn_norm.weight", "mlp.weight",
                                                             "pe.enc_content_d", "pe.weights_token", "pe.enc_token.0")):
            keep.append(n)
        elif ".merge." in n or "conv_concat" in n or "norm_concat" in n or n.startswith("extra_heads."):
            keep.append(n)
    return keep


def case_inputs(name, batch, dims):
    return ((det_tensor((batch, 1, *dims), _salt(name + ".x")) + 1.0) * 0.5).float()      # CT-like values in [0, 1)


def upstream_grad(name, key, shape):
    n = int(np.prod(shape))
    return (det_tensor(shape, _salt(name + ".go." + key)) / n ** 0.5).float()


def run_case(Model, name, mode, dims, batch, dec_prompt):
    conf = model_conf(mode, dec_prompt=dec_prompt)
    model = Model(conf).double().train()
    fill_params_deterministic(model)
    # both sides start from float32-representable values
    with torch.no_grad():
        for p in model.parameters():
            p.copy_(p.float().double())
    x = case_inputs(name, batch, dims).double()
    res = {"meta": np.array([batch, *dims, int(dec_prompt)], dtype=np.int64)}
    out = model(x)
    loss = 0.0
    for k, t in output_items(out):
        res["out." + k] = subsample(t.detach()).float().numpy()
        res["shape." + k] = np.array(t.shape, dtype=np.int64)
        loss = loss + (t * upstream_grad(name, k, t.shape).double()).sum()
    loss.backward()
    prm = dict(model.named_parameters())
    for n in grad_param_names(model):
        p = prm[n]
        if p.requires_grad:
            res["grad." + n] = (p.grad if p.grad is not None else torch.zeros_like(p)).float().numpy()
    res["trainable"] = np.array(sorted(n for n, p in prm.items() if p.requires_grad), dtype="U")
    res.update(reference_bf16_scores(Model, name, mode, dims, batch, dec_prompt, res))
    # the reference's state-dict keys and shapes: reference checkpoints must load into the mirror unchanged
    res["sd_keys"] = np.array([f"{k}:{'x'.join(str(v) for v in t.shape)}" for k, t in model.state_dict().items()], dtype="U")
    return res


def grad_error(g, ref, gmax):
    """Normalised L-inf of a gradient; an analytically-zero gradient (a conv bias in front of a Batch/InstanceNorm: 1e-17
    in float64) is measured on the scale of the largest stored gradient instead of its own round-off."""
    ref = torch.as_tensor(ref).double()
    rm = float(ref.abs().max())
    den = rm if rm > 1e-6 * gmax else gmax
    return float((g.detach().double().cpu() - ref).abs().max()) / den


def reference_bf16_scores(Model, name, mode, dims, batch, dec_prompt, res):
    """How far the REFERENCE ITSELF lands from its own float64 results when it is run the way a bf16 user runs it
    (fp32 parameters, torch.autocast(bfloat16)), on the metric of the tests.  At these toy sizes the stored gradients
    (random-projection loss through 6-12 blocks and batch statistics over a few hundred voxels) move by 5-25 %: the
    bf16 model-level test therefore bounds our deviation by max(4e-2, 1.5 x this score) per tensor."""
    model = Model(model_conf(mode, dec_prompt=dec_prompt)).double().train()
    fill_params_deterministic(model)
    model = model.float()
    x = case_inputs(name, batch, dims)
    with torch.autocast('cpu', dtype=torch.bfloat16):
        out = model(x)
    loss, scores = 0.0, {}
    for k, t in output_items(out):
        ref = torch.from_numpy(res["out." + k]).double()
        scores["bf16ref.out." + k] = np.float64(float((subsample(t.detach().float()).double() - ref).abs().max() / ref.abs().max()))
        loss = loss + (t.float() * upstream_grad(name, k, t.shape)).sum()
    loss.backward()
    prm = dict(model.named_parameters())
    gmax = max(float(np.abs(v).max()) for k, v in res.items() if k.startswith("grad."))
    for k in [k for k in res if k.startswith("grad.")]:
        scores["bf16ref." + k] = np.float64(grad_error(prm[k[5:]].grad, res[k], gmax))
    return scores


def main():
    Model = load_reference_model_class()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for i, (name, mode, dims, batch, dec_prompt) in enumerate(CASES):
        data = run_case(Model, name, mode, dims, batch, dec_prompt)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **data)
        print("wrote", name, f"{os.path.getsize(path) / 1e6:.2f} MB", [k for k in data if k.startswith("out.")])


if __name__ == "__main__":
    main()
