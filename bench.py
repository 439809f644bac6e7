"""bench.py — FlexAM Wan2.2-Fun-5B denoising-step throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one CFG-guided transformer evaluation (both branches, batch of 2) of the 30-layer FlexAM DiT on the
97-frame 512x896 latent grid (25x32x56 -> 11,200 video tokens + 448 reference tokens), synthetic inputs and
random-init weights of that architecture. Prints ONE JSON line (rank 0).

  value    steps/s with inputs resident in HBM, step-invariant caches DISABLED (every step does all the work)
  e2e      the same step through the public module call with pinned HOST inputs copied in and the prediction
           copied back inside the timed region
  roofline GEMM family (dominant: ~2/3 of the FLOPs): algorithmic FLOPs / CUDA-event time of those launches inside
           the timed region, against MEASURED_PEAKS.json's sustained bf16 figure
  cpu_baseline  the oracle (CPU port of the reference forward, fp32) on a bounded sample, extrapolated to a step
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID = (25, 32, 56)          # latent frames, rows, cols for 97 x 512 x 896
PROMPT_LENS = (37, 120)
METRIC = "denoise_steps_per_s"
UNIT = "steps/s"
WORKLOAD = "FlexAM Wan2.2-5B one denoising step, 97 frames 512x896 (11,200 tokens + 448 ref), bf16, CFG batch 2"


def real_cfg():
    return dict(model_type="ti2v", patch_size=(1, 2, 2), text_len=512, in_dim=148, dim=3072, ffn_dim=14336,
                freq_dim=256, text_dim=4096, out_dim=48, num_heads=24, num_layers=30, eps=1e-6, add_ref_conv=True,
                in_dim_ref_conv=48, add_cnn_block=True, in_dim_cnn_block=288, out_dim_cnn_block=48)


def step_flops(cfg, B=2):
    D, Fd, Lc = cfg["dim"], cfg["ffn_dim"], cfg["text_len"]
    F, H, W = GRID
    L0, R = F * (H // 2) * (W // 2), (H // 2) * (W // 2)
    L = L0 + R
    per_layer = 2 * L * D * D * 4 + 4 * L * L * D + 2 * L * D * D * 2 + 2 * Lc * D * D * 2 + 4 * L * Lc * D + 2 * L * D * Fd * 2
    front = 2 * L0 * D * cfg["in_dim"] * 4 + 2 * R * D * 48 * 4 + 2 * L * D * 192 + 2 * Lc * (cfg["text_dim"] * D + D * D)
    npix = F * H * W
    cnn = 2 * npix * 9 * (288 * 192 + 192 * 192 + 192 * 96 + 96 * 96) + 2 * npix * 96 * 48
    return B * float(cfg["num_layers"] * per_layer + front + cnn)


def ncu_traffic():
    """DRAM bytes per launch of the largest GEMM of the step (ffn.0, FX_EPI_GELU_BF16) from the committed
    `ncu --set full` capture (profiles/traffic.json, written from profiles/summary_*.md), next to its algorithmic
    bytes; None when the file is absent."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["bf16_tflops_sustained"], d["bf16_tflops"], d["hbm_gbs"], "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [ln.strip().split(", ") for ln in open(self.f.name) if ln.strip()]
        os.unlink(self.f.name)
        sm, smax, reasons, power = [], 0.0, set(), 0.0
        for r in rows:
            try:
                sm.append(float(r[0])); smax = max(smax, float(r[1])); power = max(power, float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.strip() == "Active":
                    reasons.add(name)
        sm.sort()
        # median over the samples taken under load (upper half of the clock-sorted list ~ busy samples)
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": smax or None, "power_w_max": power or None, "samples": len(sm),
                "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------
def make_host_inputs(torch, cfg, B=2, seed=1234):
    """Synthetic step inputs in pinned host memory: N(0,1) latents/controls, {0,1} mask channels with
    the first latent frame pinned, per-token timesteps (0 on first-frame tokens, 875 elsewhere), density 0.1."""
    F, H, W = GRID
    g = torch.Generator().manual_seed(seed)
    C = cfg["out_dim"]

    def pin(t):
        return t.pin_memory() if torch.cuda.is_available() else t

    x = torch.randn(B, C, F, H, W, generator=g).bfloat16()
    y = torch.randn(B, cfg["in_dim"] - C, F, H, W, generator=g)
    mask = torch.ones(F, H, W)
    mask[0] = 0
    y[:, C:C + 4] = mask
    add = torch.randn(B, cfg["in_dim_cnn_block"] - C, F, H, W, generator=g).bfloat16()
    ref = torch.randn(B, C, H, W, generator=g).bfloat16()
    ctx = [torch.randn(PROMPT_LENS[i % 2], cfg["text_dim"], generator=g).bfloat16() for i in range(B)]
    t = (mask[:, ::2, ::2].reshape(-1) * 875.0).expand(B, -1).contiguous()
    dens = torch.full((B,), 0.1)
    return dict(x=pin(x), y=pin(y.bfloat16()), additional_control=pin(add), full_ref=pin(ref),
                context=[pin(c) for c in ctx], t=pin(t), density=pin(dens), seq_len=F * (H // 2) * (W // 2))


def init_weights(torch, model, seed=1234):
    """Random init of the reference architecture on the device: Linear/conv ~ N(0, 1/fan_in), biases N(0, 0.02),
    norm weights ~ 1, modulation ~ N(0,1)/sqrt(D) (the reference's zero-init head/density MLPs are re-randomised so
    no stage is skipped numerically)."""
    g = torch.Generator(device=next(model.parameters()).device).manual_seed(seed)
    for name, p in model.named_parameters():
        if name.endswith("norm_q.weight") or name.endswith("norm_k.weight") or name.endswith("norm3.weight") or \
                (name.startswith("cnn_conv") and ".1.weight" in name):
            p.data.fill_(1.0)
        elif "modulation" in name:
            p.data.copy_((torch.randn(p.shape, device=p.device, generator=g) / math.sqrt(p.shape[-1])).to(p.dtype))
        elif name.endswith(".bias"):
            p.data.copy_((torch.randn(p.shape, device=p.device, generator=g) * 0.02).to(p.dtype))
        else:
            fan_in = p[0].numel()
            p.data.copy_((torch.randn(p.shape, device=p.device, generator=g) * fan_in ** -0.5).to(p.dtype))


def run_native(args):
    import torch
    import torch.distributed as dist
    from flexam_b200 import lib
    from flexam_b200.model import Wan2_2Transformer3DModel_FlexAM

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib.check(lib.load().fx_check_device(local), "fx_check_device")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = real_cfg()
    model = Wan2_2Transformer3DModel_FlexAM(**cfg, device=dev)
    init_weights(torch, model)
    layout = "single"
    if world > 1:
        from flexam_b200 import dist as fdist
        layout = fdist.setup(model, world, rank)
    eng = model.engine()
    host = make_host_inputs(torch, cfg)

    def to_dev(h):
        d = {k: (v.to(dev, non_blocking=True) if hasattr(v, "to") else v) for k, v in h.items() if k != "context"}
        d["context"] = [c.to(dev, non_blocking=True) for c in h["context"]]
        return d

    def call(d):
        return model(x=d["x"], t=d["t"], context=d["context"], seq_len=d["seq_len"], y=d["y"], full_ref=d["full_ref"],
                     additional_control=d["additional_control"], density=d["density"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ---------------- device-resident region (value): caches off, so no step reuses another step's work -------------
    resident = to_dev(host)
    eng.cache_static = False
    for _ in range(args.warmup):
        call(resident)
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    eng.timing = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        out = call(resident)
    ev1.record()
    barrier()
    ms_value = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
    # host time to enqueue ONE step into an empty stream (outside the timed region; inside it the launch queue is
    # full and the host is throttled by the GPU): what bounds the step once the kernels get short (8 GPUs)
    saved_timing, eng.timing = eng.timing, None
    t_host0 = time.perf_counter()
    call(resident)
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3
    barrier()
    eng.timing = saved_timing
    timing, eng.timing = eng.timing, None
    launches = eng.launches * args.steps
    fam = {}
    for kind, flops, a, b in timing:
        f = fam.setdefault(kind, [0.0, 0.0, 0])
        f[0] += flops; f[1] += a.elapsed_time(b); f[2] += 1

    if args.quick:   # profiling runs (ncu) only need the resident region
        if rank == 0:
            print(json.dumps({"quick": True, "workload": WORKLOAD, "n_gpus": world, "ms_per_step": ms_value,
                              "step_tflops": step_flops(cfg) / 1e12, "gpu_launches": launches,
                              "host_enqueue_ms_per_step": host_enqueue_ms}))
        if clocks:
            clocks.stop()
        return
    # ---------------- hoisted region (informational): step-invariant control/context work cached, as in the sampler --
    eng.cache_static = True
    call(resident)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        call(resident)
    ev1.record()
    barrier()
    ms_hoisted = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)

    # ---------------- end-to-end region: pinned host -> device copies and device -> host result every step ----------
    eng.cache_static = False    # like `value`: every end-to-end step does all the work, nothing is reused across steps
    out_host = torch.empty(out.shape, dtype=out.dtype).pin_memory()
    h2d = sum(v.numel() * v.element_size() for k, v in host.items() if hasattr(v, "numel")) + \
        sum(c.numel() * c.element_size() for c in host["context"])
    d2h = out_host.numel() * out_host.element_size()
    for _ in range(2):
        out_host.copy_(call(to_dev(host)), non_blocking=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        out_host.copy_(call(to_dev(host)), non_blocking=True)
    ev1.record()
    barrier()
    ms_e2e = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
    clk = clocks.stop() if clocks else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    sustained, burst, hbm, src = measured_peaks()
    F, H, W = GRID
    L0 = F * (H // 2) * (W // 2)
    flops = step_flops(cfg)
    gem = fam.get("gemm", [0.0, 1e-9, 0])
    att = fam.get("fmha", [0.0, 1e-9, 0])
    gemm_tfs = gem[0] / (gem[1] * 1e-3) / 1e12
    att_tfs = att[0] / (att[1] * 1e-3) / 1e12
    traffic = ncu_traffic()
    line = {
        "metric": METRIC, "value": 1e3 / ms_value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_value, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic (seeded N(0,1) latents/controls, random-init weights)",
        "config": {"workload": WORKLOAD, "layout": layout, "l2": "working set (10 GB weights + >1 GB activations per "
                   "step) exceeds the 126 MB L2; no flush needed", "timesteps": "per-token, 2 distinct values"},
        "latent_tokens_per_s": L0 * 1e3 / ms_value,
        "step_tflops": flops / 1e12, "step_frac_of_bf16_sustained": flops / (ms_value * 1e-3) / 1e12 / (world * sustained),
        "hoisted": {"ms_per_step": ms_hoisted, "value": 1e3 / ms_hoisted,
                    "note": "control fuser, text embedding and cross K/V cached across steps as in the 50-step sampler"},
        "e2e": {"value": 1e3 / ms_e2e, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": launches, "host_enqueue_ms_per_step": host_enqueue_ms,
        "roofline": {"kernel": "gemm_bf16_kernel (tcgen05, all epilogues)", "bound": "tensor", "achieved": gemm_tfs,
                     "peak": sustained, "unit": "TFLOP/s", "frac": gemm_tfs / sustained,
                     "traffic": traffic["dram_bytes_per_launch"] if traffic else None, "traffic_detail": traffic,
                     "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({src}); burst {burst}",
                     "launches": gem[2], "share_of_step": gem[1] / args.steps / ms_value},
        "roofline_fmha": {"kernel": "fmha2_fwd_kernel (tcgen05)", "bound": "tensor", "achieved": att_tfs,
                          "peak": sustained, "unit": "TFLOP/s", "frac": att_tfs / sustained, "launches": att[2],
                          "share_of_step": att[1] / args.steps / ms_value},
        "clocks": clk,
    }
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_sample(cfg, threads=None)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------
def cpu_sample(cfg, threads=None, reps=1):
    """The oracle (oracle/flexam_oracle.py, the CPU port of the reference forward) timed on a bounded sample of the
    step: ONE transformer block for ONE CFG sample at the full 11,648 tokens, fp32, all host threads; a step is 30
    blocks x 2 samples (+ <1 % front/back end), so steps/s = 1 / (60 x t_block)."""
    import torch
    from oracle import flexam_oracle as O
    n = threads or os.cpu_count() or 1
    torch.set_num_threads(n)
    ocfg = dict(dim=cfg["dim"], ffn_dim=cfg["ffn_dim"], num_heads=cfg["num_heads"], num_layers=1, eps=cfg["eps"],
                text_len=cfg["text_len"])
    F, H, W = GRID
    L = (F + 1) * (H // 2) * (W // 2)
    D, Fd = cfg["dim"], cfg["ffn_dim"]
    g = torch.Generator().manual_seed(0)
    sd = {}

    def lin(p, o, i):
        sd[p + ".weight"] = torch.randn(o, i, generator=g) * i ** -0.5
        sd[p + ".bias"] = torch.randn(o, generator=g) * 0.02
    for att in ("self_attn", "cross_attn"):
        for nm in "qkvo":
            lin(f"blocks.0.{att}.{nm}", D, D)
        sd[f"blocks.0.{att}.norm_q.weight"] = torch.ones(D)
        sd[f"blocks.0.{att}.norm_k.weight"] = torch.ones(D)
    sd["blocks.0.norm3.weight"], sd["blocks.0.norm3.bias"] = torch.ones(D), torch.zeros(D)
    lin("blocks.0.ffn.0", Fd, D), lin("blocks.0.ffn.2", D, Fd)
    sd["blocks.0.modulation"] = torch.randn(1, 6, D, generator=g) / math.sqrt(D)
    sd["blocks.0.modulation_density"] = torch.randn(1, 2, D, generator=g) / math.sqrt(D)
    x = torch.randn(L, D, generator=g)
    e0 = torch.randn(L, 6, D, generator=g) * 0.1
    dens0 = torch.randn(2, D, generator=g) * 0.1
    ctx = torch.randn(cfg["text_len"], D, generator=g)
    ang = O.rope_angles(128)
    best = float("inf")
    with torch.no_grad():
        for _ in range(reps):
            t0 = time.perf_counter()
            O.block_forward(sd, ocfg, 0, x, e0, dens0, (F + 1, H // 2, W // 2), ang, ctx, "fp32")
            best = min(best, time.perf_counter() - t0)
    step_s = best * cfg["num_layers"] * 2
    return {"value": 1.0 / step_s, "unit": UNIT, "cores": n, "kind": "port",
            "sample": f"1 of 30 blocks x 1 of 2 CFG samples at the full 11,648 tokens, fp32 oracle, {best:.2f} s; "
                      f"step extrapolated x60 = {step_s:.1f} s"}


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path is Python/torch and does not travel to the
    GPU box, so this times the oracle port on the host cores (all threads), per step a bounded sample (see cpu_sample)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = real_cfg()
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_sample(cfg)
        if i >= args.warmup:
            vals.append(r)
        if i >= 1 and (i + 1 - args.warmup) >= 2:   # keep the whole run within a few minutes
            break
    v = sum(x["value"] for x in vals) / len(vals)
    cb = dict(vals[-1], value=v)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": len(vals), "warmup": min(args.warmup, 1), "ms_per_step": 1e3 / v, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU port of the reference forward (oracle), bounded sample per step"},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="resident region only (for ncu runs)")
    ap.add_argument("--workload", default="config2", choices=["config2", "long"],
                    help="config2 = the metric's workload (default); long = BASELINE config 5, 193 frames 704x1280 "
                         "(43,120 tokens + 880 ref): a parity/stress case, not the bench line")
    a = ap.parse_args()
    if a.workload == "long":
        GRID = (49, 44, 80)
        WORKLOAD = ("FlexAM Wan2.2-5B one denoising step, 193 frames 704x1280 (43,120 tokens + 880 ref), bf16, "
                    "CFG batch 2 (long-clip stress case, BASELINE config 5)")
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
