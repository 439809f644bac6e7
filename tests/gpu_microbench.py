"""Kernel-level timing at the config-2 shapes (run on the GPU box): CUDA events, warm-up, L2 flushed between
repetitions by the working set itself or an explicit 256 MB write. Not a pytest file; prints one JSON line per case.

    python tests/gpu_microbench.py [fmha] [gemm] [rows]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flexam_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, reps=10, warm=3, flush=True):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(reps):
        if flush:
            flush_buf.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        total += a.elapsed_time(b)
    return total / reps


def bench_fmha():
    B, H, L = 2, 24, 11648
    g = torch.Generator(device=dev).manual_seed(0)
    qkv = torch.randn(B * L, 3 * H * 128, device=dev, generator=g).bfloat16()
    v5 = qkv.view(B, L, 3, H, 128)
    out = torch.empty(B, L, H, 128, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.fmha(v5[:, :, 0], v5[:, :, 1], v5[:, :, 2], out, 128 ** -0.5))
    fl = 4.0 * B * H * L * L * 128
    torch.cuda.synchronize()
    i16 = out.view(torch.int16).to(torch.int64).view(-1)
    checksum = int((i16 * (torch.arange(i16.numel(), device=dev) % 8191 + 1)).sum().item())  # same for every FX_FMHA_PIPE
    print(json.dumps({"case": "fmha_self", "poly": os.environ.get("FX_FMHA_POLY", "default"),
                      "pipe": os.environ.get("FX_FMHA_PIPE", "default"), "token": os.environ.get("FX_FMHA_TOKEN", "default"),
                      "ms": ms, "tflops": fl / ms / 1e9, "checksum": checksum}))
    # accuracy of the exp2 split against fp32 softmax on a slice
    q, k, v = v5[:1, :512, 0, :2].float(), v5[:1, :, 1, :2].float(), v5[:1, :, 2, :2].float()
    s = torch.einsum("bqhd,bkhd->bhqk", q, k) * 128 ** -0.5
    want = torch.einsum("bhqk,bkhd->bqhd", s.softmax(-1), v)
    got = out[:1, :512, :2].float()
    print(json.dumps({"case": "fmha_self_err", "rel": ((got - want).norm() / want.norm()).item()}))
    # the reference's library attention on the same tensors (F.scaled_dot_product_attention: cuDNN / flash backend),
    # timed the same way: what `attention()` (attention_utils.py:174-233) dispatches to without flash-attn
    qt, kt, vt = (v5[:, :, i].transpose(1, 2) for i in range(3))
    try:
        ms_lib = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qt, kt, vt))
        print(json.dumps({"case": "sdpa_self (library)", "ms": ms_lib, "tflops": fl / ms_lib / 1e9}))
        # steady state (1 s back to back, no flush): what both sustain under the power cap
        for name, fn in (("fmha_self", lambda: ops.fmha(v5[:, :, 0], v5[:, :, 1], v5[:, :, 2], out, 128 ** -0.5)),
                         ("sdpa_self (library)", lambda: torch.nn.functional.scaled_dot_product_attention(qt, kt, vt))):
            n = 400
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(20):
                fn()
            a.record()
            for _ in range(n):
                fn()
            b.record()
            torch.cuda.synchronize()
            ms1 = a.elapsed_time(b) / n
            print(json.dumps({"case": name + " sustained x400", "ms": ms1, "tflops": fl / ms1 / 1e9}))
    except Exception as exc:  # noqa: BLE001
        print(json.dumps({"case": "sdpa_self (library)", "error": str(exc)[:200]}))
    kv = torch.randn(B * 512, 2 * H * 128, device=dev, generator=g).bfloat16().view(B, 512, 2, H, 128)
    ms = timeit(lambda: ops.fmha(v5[:, :, 0], kv[:, :, 0], kv[:, :, 1], out, 128 ** -0.5))
    print(json.dumps({"case": "fmha_cross", "ms": ms, "tflops": 4.0 * B * H * L * 512 * 128 / ms / 1e9}))


def bench_gemm():
    M = int(os.environ.get("FX_BENCH_M", 2 * 11648))   # 2912 = rows per rank at 8 GPUs (cfg2 x sp4)
    g = torch.Generator(device=dev).manual_seed(0)
    for name, N, K, epi in (("qkv", 9216, 3072, 0), ("o_resid", 3072, 3072, 3), ("cq", 3072, 3072, 0),
                            ("ffn1_gelu", 14336, 3072, 1), ("ffn2_resid", 3072, 14336, 3)):
        a = (torch.randn(M, K, device=dev, generator=g) * 0.5).bfloat16()
        w = (torch.randn(N, K, device=dev, generator=g) * K ** -0.5).bfloat16()
        b = torch.randn(N, device=dev, generator=g).bfloat16()
        out = torch.zeros(M, N, device=dev, dtype=torch.float32 if epi == 3 else torch.bfloat16)
        gm = torch.randn(N, device=dev, generator=g)
        ge = torch.randn(2, 6, N, device=dev, generator=g)
        idx = torch.randint(0, 2, (M,), device=dev, generator=g, dtype=torch.int32)
        if epi == 3:
            fn = lambda: ops.gemm(a, w, b, out, 3, gate_mod=gm, gate_e=ge[:, 2], row_idx=idx)  # noqa: E731
        else:
            fn = lambda: ops.gemm(a, w, b, out, epi)  # noqa: E731
        ms = timeit(fn)
        print(json.dumps({"case": "gemm_" + name, "M": M, "mode": os.environ.get("FX_GEMM_MODE", "default"), "ms": ms,
                          "tflops": 2.0 * M * N * K / ms / 1e9}))


def bench_rows():
    M, D = 2 * 11648, 3072
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(M, D, device=dev, generator=g)
    out = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    mod = torch.randn(6, D, device=dev, generator=g)
    e = torch.randn(2, 6, D, device=dev, generator=g)
    dens = torch.randn(2, 2, D, device=dev, generator=g)
    idx = torch.randint(0, 2, (M,), device=dev, generator=g, dtype=torch.int32)
    ms = timeit(lambda: ops.ln_modulate(x, out, 1e-6, mod[0], mod[1], e[:, 0], e[:, 1], 6 * D, idx, mod[2], dens[:, 0],
                                        2 * D, M // 2))
    print(json.dumps({"case": "ln_modulate", "ms": ms, "gbs": M * D * 6 / ms / 1e6}))
    gam = torch.ones(D, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.ln_affine(x, out, 1e-6, gam, gam))
    print(json.dumps({"case": "ln_affine", "ms": ms, "gbs": M * D * 6 / ms / 1e6}))
    qkv = torch.randn(M, 3 * D, device=dev, generator=g).bfloat16()
    from flexam_b200.model import rope_table
    fr = rope_table(128).to(dev)
    ms = timeit(lambda: ops.rmsnorm_rope(qkv[:, :2 * D], gam, 1e-6, fr, (26, 16, 28), 0, M // 2, weight2=gam))
    print(json.dumps({"case": "rmsnorm_rope_qk", "ms": ms, "gbs": M * D * 8 / ms / 1e6}))
    cq = torch.randn(M, D, device=dev, generator=g).bfloat16()
    ms = timeit(lambda: ops.rmsnorm_rope(cq, gam, 1e-6))
    print(json.dumps({"case": "rmsnorm_cq", "ms": ms, "gbs": M * D * 4 / ms / 1e6}))


def bench_chain():
    """One rank's share of a transformer block at 8 GPUs (cfg2 x sp4: 2912 tokens, 6 heads over all 11,648 keys),
    repeated 30 times back to back: what FX_PDL (programmatic dependent launch) is meant to speed up."""
    M, D, Fd, H, L = int(os.environ.get("FX_BENCH_M", 2912)), 3072, 14336, 6, 11648
    g = torch.Generator(device=dev).manual_seed(0)
    rnd = lambda *sh: torch.randn(*sh, device=dev, generator=g)  # noqa: E731
    x = rnd(M, D)
    h = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    gam = torch.ones(D, device=dev, dtype=torch.bfloat16)
    wqkv, bqkv = (rnd(3 * D, D) * D ** -0.5).bfloat16(), rnd(3 * D).bfloat16()
    wo, bo = (rnd(D, D) * D ** -0.5).bfloat16(), rnd(D).bfloat16()
    w1, b1 = (rnd(Fd, D) * D ** -0.5).bfloat16(), rnd(Fd).bfloat16()
    w2, b2 = (rnd(D, Fd) * Fd ** -0.5).bfloat16(), rnd(D).bfloat16()
    qkv = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
    ffn = torch.empty(M, Fd, device=dev, dtype=torch.bfloat16)
    cq = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    aqkv = rnd(L, 3 * H * 128).bfloat16().view(1, L, 3, H, 128)       # the rank's heads over the whole sequence
    aout = torch.empty(1, L, H, 128, device=dev, dtype=torch.bfloat16)
    ckv = rnd(512, 2 * 24 * 128).bfloat16().view(1, 512, 2, 24, 128)
    cout = torch.empty(1, M, 24, 128, device=dev, dtype=torch.bfloat16)

    def block():
        ops.ln_affine(x, h, 1e-6, gam, gam)
        ops.gemm(h, wqkv, bqkv, qkv, 0)
        ops.rmsnorm_rope(qkv[:, :2 * D], gam, 1e-6, weight2=gam)
        ops.fmha(aqkv[:, :, 0], aqkv[:, :, 1], aqkv[:, :, 2], aout, 128 ** -0.5)
        ops.gemm(h, wo, bo, x, 3)
        ops.ln_affine(x, h, 1e-6, gam, gam)
        ops.gemm(h, wo, bo, cq, 0)
        ops.rmsnorm_rope(cq, gam, 1e-6)
        ops.fmha(cq.view(1, M, 24, 128), ckv[:, :, 0], ckv[:, :, 1], cout, 128 ** -0.5)
        ops.gemm(cout.view(M, D), wo, bo, x, 3)
        ops.ln_affine(x, h, 1e-6, gam, gam)
        ops.gemm(h, w1, b1, ffn, 1)
        ops.gemm(ffn, w2, b2, x, 3)

    def step():
        for _ in range(30):
            block()

    ms = timeit(step, reps=5, warm=2, flush=False)
    print(json.dumps({"case": "block_chain_x30", "M": M, "pdl": os.environ.get("FX_PDL", "0"), "ms": ms,
                      "launches": 13 * 30, "checksum": int(x.view(torch.int32).to(torch.int64).sum().item())}))


if __name__ == "__main__":
    which = sys.argv[1:] or ["fmha", "gemm", "rows"]
    if "fmha" in which:
        bench_fmha()
    if "gemm" in which:
        bench_gemm()
    if "rows" in which:
        bench_rows()
    if "chain" in which:
        bench_chain()
