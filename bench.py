#!/usr/bin/env python
"""bench.py -- throughput of the prompted 3D shifted-window attention hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dtype bf16|fp32]

One "step" = forward + backward of the hot path over one batch of synthetic input: the Swin encoder of
SwinUNETR(feature_size=48) on 96^3 patches with encoder prompting (BASELINE config[1]): three
ConsecutiveSwinBlocks stages = 6 prompted window-attention blocks (+ the 3 PatchMerging layers that chain
them) on the patch-embedded feature map [B, 48, 48, 48, 48].  Metric: 3D patches / s, whole job.

Prints ONE JSON line (rank 0).  Under torchrun (N > 1) every rank processes its own batch shard (weak
scaling: fixed per-GPU batch) and gradients are all-reduced over NCCL once per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WS = (8, 8, 4)
E = 64
I_PROMPT = 64
FEATURE = 48
HEADS_ENC = 4
PATCH = 96
METRIC = "3D patches/sec fwd+bwd (96^3 SwinUNETR+prompts) at 1/2/4/8 B200; attn TFLOP/s"
UNIT = "patches/s"


def stage_specs(patch=PATCH, feature=FEATURE):
    """(C, heads, dims) of the three encoder stages after the 2x2x2 patch embedding (swin_unetr.py:146-178)."""
    d = patch // 2
    return [(feature, HEADS_ENC, (d, d, d)), (2 * feature, 2 * HEADS_ENC, (d // 2, d // 2, d // 2)),
            (4 * feature, 4 * HEADS_ENC, (d // 4, d // 4, d // 2))]


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback")


# --------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def build_encoder(device, use_checkpoint=False, attn_drop=0.0, proj_drop=0.0):
    import pwa_b200
    torch.manual_seed(0)
    stages, prompts = [], []
    for i, (c, h, _) in enumerate(stage_specs()):
        stages.append(pwa_b200.ConsecutiveSwinBlocks(hidden_channels=c, num_heads=h, pos_bias_embed_dim=E, max_prompts=1,
                                                     tokens_per_prompt=I_PROMPT, window_size=WS, use_token_params=True,
                                                     down=True, merge_last_dim=(i < 1), use_checkpoint=use_checkpoint,
                                                     attn_drop=attn_drop, proj_drop=proj_drop))
        for _ in range(2):   # prompt_tokens['enc'][2j], [2j+1]  (swin_unetr.py:400-409), xavier-uniform
            prompts.append(torch.nn.Parameter(torch.nn.init.xavier_uniform_(torch.empty(I_PROMPT, c))))
    model = torch.nn.ModuleList(stages).to(device)
    plist = torch.nn.ParameterList(prompts).to(device)
    return model, plist


class _MeanSquare(torch.autograd.Function):
    """mean(x.float() ** 2), the stand-in loss on every stage output, as ONE reduction kernel forward and ONE scaling
    kernel backward.  The plain torch expression costs three + six elementwise passes over every stage output (fp32
    copy, pow, mean; fill, three muls, two copies): ~4 % of the step spent outside the path this bench measures."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        n = torch.linalg.vector_norm(x, 2, dtype=torch.float32)
        return n * n / x.numel()

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        # (a 0-dim CUDA operand sends `x * coef` down TensorIterator's strided path, 21 us at the first stage; the
        #  multi-tensor kernel with a tensor scalar is vectorised)
        coef = (g * (2.0 / x.numel())).to(x.dtype)
        return torch._foreach_mul((x,), coef)[0]


def encoder_step(model, plist, x):
    """forward + backward; returns the scalar loss tensor.  Prompts are broadcast as in
    swin_unetr.py:56-60 (.unsqueeze(0).repeat(B,1,1))."""
    b = x.shape[0]
    loss = 0.0
    for j, stage in enumerate(model):
        p_w = plist[2 * j].to(x.dtype).unsqueeze(0).expand(b, -1, -1)
        p_sw = plist[2 * j + 1].to(x.dtype).unsqueeze(0).expand(b, -1, -1)
        x = stage(x, (p_w, p_sw))
        loss = loss + _MeanSquare.apply(x)          # every stage output feeds the decoder/heads in the real model
    loss.backward()
    return loss.detach()


def allreduce_grads(params, world):
    """One flat-bucket NCCL all-reduce of all gradients (sum -> / world)."""
    import torch.distributed as dist
    grads = [p.grad for p in params if p.grad is not None]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat)
    flat.div_(world)
    o = 0
    for g in grads:
        g.copy_(flat[o:o + g.numel()].view_as(g))
        o += g.numel()


def algorithmic_work(B):
    """Per step: attention FLOPs (12*B*P*N*N'*C fwd+bwd) and partition/reverse bytes."""
    import pwa_b200
    n = WS[0] * WS[1] * WS[2]
    flops = 0.0
    for c, h, dims in stage_specs():
        g = pwa_b200.get_geometry(dims, WS, (0, 0, 0))
        flops += 2 * 12.0 * B * g.P * n * (n + I_PROMPT) * c
    return flops


def run_ours(args):
    import torch.distributed as dist
    import pwa_b200
    from pwa_b200.functional import KernelStats

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:
        emit(json.dumps({"error": "launch with torchrun for --gpus > 1"}))
        return 2
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    B = args.batch
    model, plist = build_encoder(dev, use_checkpoint=args.checkpoint, attn_drop=args.dropout,
                                 proj_drop=0.0 if args.checkpoint else args.dropout)
    params = list(model.parameters()) + list(plist.parameters())
    c0, _, d0 = stage_specs()[0]
    gen = torch.Generator().manual_seed(1234 + rank)
    # several distinct pinned host batches; every step's working set (inputs + saved activations, > 1 GB) is
    # far larger than the 126 MB L2, so no explicit flush is needed between timed iterations
    host = [torch.randn(B, c0, *d0, generator=gen).to(dtype).pin_memory() for _ in range(2)]
    xdev = [h.to(dev) for h in host]

    def zero():
        for p in params:
            p.grad = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The whole step (forward + loss + backward through the module API) is captured once into a CUDA graph and
    # replayed: ~450 short kernels per step are otherwise bound by Python/launch overhead (pwa_b200/graphs.py).
    graphed = None
    if not args.no_graph:
        from pwa_b200.graphs import GraphedStep
        graphed = GraphedStep(lambda x: encoder_step(model, plist, x), [xdev[0].clone().requires_grad_(True)], params,
                              flat_grads=world > 1)

    def run_step(x_src):
        if graphed is not None:
            loss = graphed(x_src)                     # copy into the static input (H2D or D2D) + one graph launch
            if world > 1:
                graphed.allreduce_flat()              # ONE in-place NCCL all-reduce of the flat gradient buffer
        else:
            zero()
            loss = encoder_step(model, plist, x_src.to(dev, non_blocking=True).clone().requires_grad_(True))
            if world > 1:
                allreduce_grads(params, world)
        return loss

    def step_resident(i):
        return run_step(xdev[i % len(xdev)])

    from pwa_b200.graphs import InputPrefetcher
    feeder = InputPrefetcher(xdev[0], dev)

    loss_host = torch.empty(2, dtype=torch.float32).pin_memory()
    loss_events = [torch.cuda.Event(), torch.cuda.Event()]

    def run_e2e(n):
        """n steps from pinned host batches: the H2D copy of step i+1 runs on a side stream while step i computes.
        The loss of EVERY step is read on the host, one step late (as a training loop logs it): step i's loss goes to
        pinned host memory with an asynchronous D2H copy ordered behind its graph replay, and is waited for and read
        after step i+1 has been launched, so that the host never drains the device between steps; the last one is
        read before the function returns (inside the timed region)."""
        feeder.prefetch(host[0])
        total = 0.0
        for i in range(n):
            x = feeder.get()
            if i + 1 < n:
                feeder.prefetch(host[(i + 1) % len(host)])
            loss = run_step(x)
            loss_host[i % 2:i % 2 + 1].copy_(loss.reshape(1), non_blocking=True)
            loss_events[i % 2].record()
            if i > 0:
                loss_events[(i - 1) % 2].synchronize()
                total += float(loss_host[(i - 1) % 2])
        loss_events[(n - 1) % 2].synchronize()
        total += float(loss_host[(n - 1) % 2])
        return total

    def step_eager_instrumented(i):
        zero()
        return encoder_step(model, plist, xdev[i % len(xdev)].clone().requires_grad_(True))

    for i in range(args.warmup):
        step_resident(i)
    barrier()

    # ---- timed region 1: inputs resident in HBM ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    KernelStats.reset(enabled=True, timing=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if args.cuda_profiler_range:
        torch.cuda.profiler.start()
    e0.record()
    for i in range(args.steps):
        step_resident(i)
    e1.record()
    barrier()
    if args.cuda_profiler_range:
        torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    launches = KernelStats.launches
    KernelStats.reset(enabled=False)

    # ---- timed region 2: end to end through the public module API with host buffers ----
    run_e2e(min(2, args.warmup))
    barrier()
    t0 = time.perf_counter()
    e0.record()
    run_e2e(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    wall_e2e = (time.perf_counter() - t0) * 1e3
    ms_e2e = max(ms_e2e, wall_e2e)
    # ---- per-kernel device times: the same steps run eagerly with CUDA events around every C-ABI call on the
    # launching stream (events cannot be recorded inside a graph replay); feeds `roofline` and `kernels` ----
    ksum = {}
    if rank == 0:
        KernelStats.reset(enabled=True, timing=True)
        for i in range(args.steps):
            step_eager_instrumented(i)
        torch.cuda.synchronize()
        ksum = KernelStats.summary()
        KernelStats.reset(enabled=False)
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_sample()

    if rank == 0:
        pk = peaks()
        value = B * world * args.steps / (ms / 1e3)
        e2e_v = B * world * args.steps / (ms_e2e / 1e3)
        # dominant kernel by accumulated device time
        dom = max(ksum.items(), key=lambda kv: kv[1][1]) if ksum else None
        roof = None
        try:   # DRAM traffic per launch of the dominant kernel, from the committed ncu --set full capture
            traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            traffic_tab = {}
        if dom is not None:
            name, (calls, tms, work) = dom
            if name.startswith("attn"):
                ach = work / (tms / 1e3) / 1e12
                roof = {"kernel": name, "bound": "tensor", "achieved": round(ach, 3), "peak": pk["tf_sus"],
                        "unit": "TFLOP/s", "frac": round(ach / pk["tf_sus"], 5),
                        "traffic": traffic_tab.get(name, {}).get("bytes"), "traffic_of": traffic_tab.get(name, {}).get("launch"),
                        "peak_source": pk["src"] + " (sustained bf16 cuBLAS)", "avg_launch_ms": round(tms / calls, 4),
                        "share_of_step": round(tms / ms, 4),
                        "timed": "CUDA events around each launch in an eager pass of the same steps"}
            else:
                ach = work / (tms / 1e3) / 1e9
                roof = {"kernel": name, "bound": "hbm", "achieved": round(ach, 1), "peak": pk["hbm"], "unit": "GB/s",
                        "frac": round(ach / pk["hbm"], 5), "traffic": traffic_tab.get(name, {}).get("bytes"),
                        "traffic_of": traffic_tab.get(name, {}).get("launch"), "peak_source": pk["src"],
                        "avg_launch_ms": round(tms / calls, 4), "share_of_step": round(tms / ms, 4)}
        kern = {}
        for name, (calls, tms, work) in ksum.items():
            unit = "TFLOP/s" if name.startswith("attn") else "GB/s"
            rate = work / (tms / 1e3) / (1e12 if unit == "TFLOP/s" else 1e9)
            peak = pk["tf_sus"] if unit == "TFLOP/s" else pk["hbm"]
            kern[name] = {"calls": calls, "ms_total": round(tms, 3), "rate": round(rate, 2), "unit": unit,
                          "frac_of_peak": round(rate / peak, 4)}
        line = {
            "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"SwinUNETR(feature_size=48) encoder hot path, 96^3 patches: 3 ConsecutiveSwinBlocks "
                                   f"stages = 6 prompted window-attention blocks (ws 8x8x4, 64 prompt tokens/block) + 3 "
                                   f"PatchMerging, fwd+bwd, batch {B}/GPU, {args.dtype}; random-init weights",
                       "per_gpu_batch": B, "global_batch": B * world, "patch": PATCH, "parallelism": f"dp{world}",
                       "l2": "per-step working set (>1 GB of activations) exceeds the 126 MB L2; inputs rotate",
                       "launch": "eager" if graphed is None else "whole step (fwd+loss+bwd) replayed as one CUDA graph",
                       "dropout": args.dropout, "use_checkpoint": bool(args.checkpoint)},
            "e2e": {"value": round(e2e_v, 4), "unit": UNIT,
                    "h2d_bytes_per_step": host[0].numel() * host[0].element_size(), "d2h_bytes_per_step": 4},
            "gpu_launches": launches,
            "attn_tflops": round(algorithmic_work(B * world) * args.steps / (ms / 1e3) / 1e12, 3),
            "roofline": roof, "kernels": kern, "clocks": clocks, "cpu_baseline": cpu,
        }
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


# --------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's algorithm for this path on the host cores
# --------------------------------------------------------------------------------------------------
def _oracle_encoder(threads):
    """The oracle restatement (validated against the live reference's golden vectors; the reference tree
    itself does not exist on the GPU box) of the same 3-stage encoder, fp32 on CPU."""
    from oracle import restatement as R
    import pwa_b200  # only for module construction (same init as our arm); compute is the oracle's
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    sds, prompts = [], []
    for i, (c, h, _) in enumerate(stage_specs()):
        m = pwa_b200.ConsecutiveSwinBlocks(hidden_channels=c, num_heads=h, pos_bias_embed_dim=E, max_prompts=1,
                                           tokens_per_prompt=I_PROMPT, window_size=WS, down=True, merge_last_dim=(i < 1))
        sds.append({k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in m.state_dict().items()})
        prompts += [torch.nn.init.xavier_uniform_(torch.empty(I_PROMPT, c)).requires_grad_(True) for _ in range(2)]

    def step(x):
        loss = 0.0
        for j, (c, h, _) in enumerate(stage_specs()):
            pw = prompts[2 * j].unsqueeze(0).repeat(x.shape[0], 1, 1)
            ps = prompts[2 * j + 1].unsqueeze(0).repeat(x.shape[0], 1, 1)
            x = R.pair_forward(sds[j], x, (pw, ps), WS, h, True, j < 1)
            loss = loss + x.square().mean()
        loss.backward()
        return float(loss.detach())
    return step


def cpu_baseline_sample(steps=8, warmup=1, batch=1):
    threads = os.cpu_count() or 1
    step = _oracle_encoder(threads)
    c0, _, d0 = stage_specs()[0]
    gen = torch.Generator().manual_seed(1234)
    x = torch.randn(batch, c0, *d0, generator=gen)
    for _ in range(warmup):
        step(x.clone().requires_grad_(True))
    t0 = time.perf_counter()
    for _ in range(steps):
        step(x.clone().requires_grad_(True))
    dt = time.perf_counter() - t0
    return {"value": round(batch * steps / dt, 5), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{steps} step(s) of the same 6-block encoder workload at batch {batch}, fp32, oracle restatement "
                      f"(torch CPU, {threads} threads), {dt:.1f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps, warmup = max(1, min(args.steps, 8)), max(1, min(args.warmup, 1))
    cpu = cpu_baseline_sample(steps=steps, warmup=warmup, batch=1)
    line = {
        "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": round(1e3 / cpu["value"], 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "same 6-block SwinUNETR encoder hot path, 96^3, fwd+bwd; reference algorithm on the host "
                               "CPU (oracle port, batch 1 per step)", "per_gpu_batch": 1, "patch": PATCH},
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(json.dumps(line))
    return 0


_JSON_FD = None


def _claim_stdout():
    """stdout carries exactly ONE JSON line.  Libraries print there too (NCCL's version banner, through C stdio, whatever
    NCCL_DEBUG_FILE says), so file descriptor 1 is pointed at stderr for the whole run and the result line is written
    to a private duplicate of the original stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: str):
    if _JSON_FD is None:
        print(line, flush=True)
    else:
        os.write(_JSON_FD, (line + "\n").encode())


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=4, help="per-GPU batch (BASELINE config[1]: 4)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dropout", type=float, default=0.0,
                    help="attn_drop = proj_drop of the blocks (the reference's example config uses 0.1; the headline number is "
                         "measured without dropout, like the parity tests)")
    ap.add_argument("--checkpoint", action="store_true",
                    help="use_checkpoint=True as in the reference's example config (activation recomputation; with --dropout only "
                         "attn_drop is applied, torch's proj_drop cannot be recomputed inside a graph capture)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--cuda-profiler-range", action="store_true",
                    help="bracket the HBM-resident timed region with cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
