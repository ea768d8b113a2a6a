"""WindowAttention.forward with the reference's literal arguments -- dense pos_bias / mask tensors or None
(window_attention.py:35-61) -- on the dense-argument kernels (csrc/attn_dense.cu), against goldens of the LIVE reference
module (tests/golden/attn_dense.npz, oracle/gen_golden.py --dense-only) and against the compact (BiasTables / region id)
form of the same attention."""
import pytest
import torch

import pwa_b200
from pwa_b200 import functional as PF
from oracle import restatement as R
from tests.util import load_npz, rel_linf

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CASES = [("block_form", 12, 4), ("bias_only", 12, 2), ("mask_only", 24, 4), ("plain", 12, 4), ("full_rank", 12, 2)]


def _load(name):
    d = load_npz("attn_dense")
    return (lambda k: torch.from_numpy(d[f"{name}.{k}"]) if f"{name}.{k}" in d else None), d


@pytest.mark.parametrize("name,C,heads", CASES)
@pytest.mark.parametrize("dtype,rtol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
def test_dense_arguments_vs_reference_golden(name, C, heads, dtype, rtol):
    """Output and the gradients of q, k, v, pos_bias and all five parameter tensors.  fp32: 1e-4 against the reference's
    float64 results.  bf16: 2e-2 against the ORACLE run in float64 on the bf16-rounded inputs and weights (the golden's
    fp32 inputs are not representable in bf16; the oracle itself is held to the golden by tests/test_oracle_golden.py)."""
    get, d = _load(name)
    mod = pwa_b200.WindowAttention(C, heads).to(DEV)
    sd = {k[len(name) + 4:]: torch.from_numpy(v) for k, v in d.items() if k.startswith(name + ".sd.")}
    mod.load_state_dict(sd)
    mod = mod.to(dtype).eval()
    q, k, v = (get(t).to(DEV, dtype).requires_grad_(True) for t in "qkv")
    bias, mask, go = get("bias"), get("mask"), get("go")
    bias_d = bias.to(DEV).requires_grad_(True) if bias is not None else None
    y = mod(q, k, v, pos_bias=bias_d, mask=None if mask is None else mask.to(DEV))
    assert y.shape == q.shape and y.dtype == dtype
    y.backward(go.to(DEV, dtype))
    if dtype == torch.float32:
        ref_out = get("out")
        ref_g = {n: get("grad." + n) for n in ("q", "k", "v", "bias")}
        ref_p = {n: get("grad." + n) for n in sd}
    else:
        sd64 = {n: t.to(dtype).double().requires_grad_(True) for n, t in sd.items()}
        l = [get(t).to(dtype).double().requires_grad_(True) for t in "qkv"]
        b64 = bias.double().requires_grad_(True) if bias is not None else None
        ref_out = R.window_attention_module(sd64, *l, b64, None if mask is None else mask.double(), heads)
        (ref_out * go.to(dtype).double()).sum().backward()
        ref_g = {"q": l[0].grad, "k": l[1].grad, "v": l[2].grad, "bias": None if b64 is None else b64.grad}
        ref_p = {n: t.grad for n, t in sd64.items()}
    assert rel_linf(y, ref_out) < rtol
    for n, t in (("q", q), ("k", k), ("v", v)):
        assert rel_linf(t.grad, ref_g[n]) < rtol, n
    if bias is not None:
        assert bias_d.grad.shape == bias.shape
        assert rel_linf(bias_d.grad, ref_g["bias"]) < rtol
    for n, prm in mod.named_parameters():
        assert rel_linf(prm.grad, ref_p[n]) < rtol, n


def test_dense_form_equals_compact_form_on_a_shifted_window_set():
    """The block's inputs expanded the way the reference does it (RelativePE.forward's dense bias, get_attn_mask's dense
    mask extended over the prompt rows, prompt rows appended to q / k / v: swin_block.py:187-214) through the dense kernels
    == the compact form (tables, region ids, prompts as separate K/V) through the fused kernels, on the content rows."""
    torch.manual_seed(3)
    B, C, heads, I, ws, dims, shift = 2, 48, 4, 64, (8, 8, 4), (16, 16, 8), (4, 4, 2)
    geom = pwa_b200.get_geometry(dims, ws, shift)
    P, N = geom.P, geom.N
    ids = geom.region_ids(DEV)
    mod = pwa_b200.WindowAttention(C, heads).to(DEV).eval()
    th, tw, td = (0.5 * torch.randn(heads, w, w, device=DEV) for w in ws)
    tok = 0.5 * torch.randn(heads, I, device=DEV)
    x = torch.randn(B, P, N, C, device=DEV)
    pr = torch.randn(B, I, C, device=DEV)
    compact = mod(x, x, x, pos_bias=pwa_b200.BiasTables(th, tw, td, tok, ws), mask=ids, prompts=pr)
    bias = R.dense_bias(th.cpu(), tw.cpu(), td.cpu(), tok.cpu()).to(DEV)                      # [h, N, N+I]
    bias_full = torch.zeros(1, heads, N + I, N + I, device=DEV)
    bias_full[0, :, :N] = bias
    m = (ids[:, :, None] == ids[:, None, :]).float()                                          # [P, N, N]
    mask_full = torch.zeros(1, P, 1, N + I, N + I, device=DEV)
    mask_full[0, :, 0, :N, :N] = m
    mask_full[0, :, 0, :N, N:] = 1.0
    xp = torch.cat([x, pr[:, None].expand(-1, P, -1, -1)], dim=2)
    dense = mod(xp, xp, xp, pos_bias=bias_full, mask=mask_full)
    assert dense.shape == (B, P, N + I, C)
    assert rel_linf(dense[:, :, :N], compact) < 1e-4


def test_dense_form_dropout_matches_oracle_mask_and_argument_errors():
    """Training-mode attention dropout in the dense form uses the generator of the fused kernels (n_q rows per
    window-head): forward and gradients against the oracle with the restated mask; malformed arguments raise."""
    torch.manual_seed(5)
    b, p, nq, nk, C, heads = 2, 3, 16, 40, 12, 4
    q, k, v = (torch.randn(b, p, n, C, device=DEV, requires_grad=True) for n in (nq, nk, nk))
    bias = torch.randn(heads, nq, nk, device=DEV, requires_grad=True)
    words = [2024, 77]
    seed = torch.tensor(words, dtype=torch.int32, device=DEV)
    out = PF.dense_window_attention(q, k, v, bias, None, heads, 3 ** -0.5, p_drop=0.25, seed=seed)
    go = torch.randn_like(out)
    out.backward(go)
    keep = R.dropout_keep_factor_torch(words, 0, b * p, heads, nq, nk, 0.25).reshape(b, p, heads, nq, nk).double()
    assert abs((keep == 0).double().mean().item() - 0.25) < 0.03
    l = [t.detach().double().cpu().requires_grad_(True) for t in (q, k, v, bias)]
    eye = torch.eye(C, dtype=torch.float64)
    sd = {"to_q.weight": eye, "to_k.weight": eye, "to_v.weight": eye, "proj.weight": eye, "proj.bias": torch.zeros(C, dtype=torch.float64)}
    ref = R.window_attention_module(sd, l[0], l[1], l[2], l[3], None, heads, drop=keep)
    (ref * go.double().cpu()).sum().backward()
    assert rel_linf(out, ref) < 1e-4
    for t, r, n in zip((q, k, v, bias), l, "qkvb"):
        assert rel_linf(t.grad, r.grad) < 1e-4, n
    mod = pwa_b200.WindowAttention(C, heads).to(DEV)
    with pytest.raises(ValueError):
        mod(q, k, v, pos_bias=torch.zeros(heads, nq + 1, nk, device=DEV))                    # does not broadcast
    with pytest.raises(ValueError):
        mod(q, k, v, pos_bias=None, mask=None, prompts=torch.zeros(b, 8, C, device=DEV))     # prompts belong to the compact form
    with pytest.raises(RuntimeError):
        mod(q.cpu(), k.cpu(), v.cpu())                                                       # no CPU fallback
