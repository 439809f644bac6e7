"""First-contact diagnostics for the sm_100a kernels (run on the GPU box, not a pytest file).

Each check runs in its own subprocess under a timeout so a deadlocked mbarrier pipeline cannot take the
whole call down; results are printed as one line per check and also written to gpurun_out/diag.json.
Usage: python tests/gpu_diag.py [check ...]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _rel(a, b):
    import torch
    a = a.float()
    b = b.float()
    return (torch.linalg.vector_norm(a - b) / (torch.linalg.vector_norm(b) + 1e-30)).item()


def chk_gemm(M, N, K, epi=0, bias=True):
    import torch
    from flexam_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    b = (torch.randn(N, device="cuda", generator=g)).bfloat16() if bias else None
    ref = a.float() @ w.float().t()
    if bias:
        ref = ref + b.float()
    ref_b = ref.bfloat16().float()
    if epi == 0:
        out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
        ops.gemm(a, w, b, out, 0)
        want = ref_b
    elif epi == 1:
        out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
        ops.gemm(a, w, b, out, 1)
        want = torch.nn.functional.gelu(ref_b, approximate="tanh")
    elif epi == 2:
        out = torch.zeros(M, N, device="cuda", dtype=torch.float32)
        ops.gemm(a, w, b, out, 2)
        want = ref_b
    else:
        x0 = torch.randn(M, N, device="cuda", generator=g)
        out = x0.clone()
        U = 3
        gate_mod = torch.randn(N, device="cuda", generator=g)
        gate_e = torch.randn(U, 6, N, device="cuda", generator=g)
        idx = torch.randint(0, U, (M,), device="cuda", generator=g, dtype=torch.int32)
        ops.gemm(a, w, b, out, 3, gate_mod=gate_mod, gate_e=gate_e[:, 2], row_idx=idx)
        want = x0 + ref_b * (gate_mod[None] + gate_e[idx.long(), 2])
    torch.cuda.synchronize()
    err = _rel(out, want)
    # error pattern by 128-row / 64-col blocks helps to localise descriptor or swizzle mistakes
    d = (out.float() - want).abs()
    info = {"rel": err, "max": d.max().item(), "row_blk_max": [round(x, 4) for x in
            d[: min(M, 512)].reshape(-1, min(128, M), N).amax(dim=(1, 2)).tolist()] if M % 128 == 0 and M >= 128 else []}
    return err < 2e-2, info


def chk_fmha(B, H, Lq, Lk):
    import torch
    from flexam_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(2)
    q = torch.randn(B, Lq, H, 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(B, Lk, H, 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(B, Lk, H, 128, device="cuda", generator=g).bfloat16()
    out = torch.zeros(B, Lq, H, 128, device="cuda", dtype=torch.bfloat16)
    ops.fmha(q, k, v, out, 128 ** -0.5)
    torch.cuda.synchronize()
    s = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float()) * 128 ** -0.5
    want = torch.einsum("bhqk,bkhd->bqhd", s.softmax(-1), v.float())
    err = _rel(out, want)
    d = (out.float() - want).abs()
    return err < 2e-2, {"rel": err, "max": d.max().item(), "nan": bool(torch.isnan(out.float()).any())}


def chk_ln():
    import torch
    from flexam_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    M, D, U, B = 1344, 3072, 3, 2
    x = torch.randn(M, D, device="cuda", generator=g) * 2 + 0.3
    mod = torch.randn(6, D, device="cuda", generator=g) * 0.1
    e = torch.randn(U, 6, D, device="cuda", generator=g) * 0.1
    dens = torch.randn(B, 2, D, device="cuda", generator=g) * 0.1
    idx = torch.randint(0, U, (M,), device="cuda", generator=g, dtype=torch.int32)
    out = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    ops.ln_modulate(x, out, 1e-6, mod[0], mod[1], e[:, 0], e[:, 1], 6 * D, idx, None, dens[:, 0], 2 * D, M // B)
    ln = torch.nn.functional.layer_norm(x, (D,), eps=1e-6)
    b = torch.arange(M, device="cuda") // (M // B)
    want = ln * (1 + mod[1] + e[idx.long(), 1]) + mod[0] + e[idx.long(), 0] + dens[b, 0]
    e1 = _rel(out, want)
    gam = torch.randn(D, device="cuda", generator=g).bfloat16()
    bet = torch.randn(D, device="cuda", generator=g).bfloat16()
    out2 = torch.empty_like(out)
    ops.ln_affine(x, out2, 1e-6, gam, bet)
    e2 = _rel(out2, ln * gam.float() + bet.float())
    return e1 < 5e-3 and e2 < 5e-3, {"mod": e1, "affine": e2}


def chk_rms():
    import torch
    from flexam_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(4)
    B, L, D = 2, 96, 3072
    gf, gh, gw = 4, 4, 6  # 96 tokens
    buf = torch.randn(B * L, 3 * D, device="cuda", generator=g).bfloat16()
    w = (1 + 0.1 * torch.randn(D, device="cuda", generator=g)).bfloat16()
    ang = torch.rand(1024, 64, device="cuda", generator=g, dtype=torch.float64) * 6.28
    freqs = torch.stack([ang.cos(), ang.sin()], -1).float().contiguous()
    x = buf[:, :D].clone()
    ops.rmsnorm_rope(buf[:, :D], w, 1e-6, freqs, (gf, gh, gw), 0, L)
    xf = x.float()
    r = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6).bfloat16()
    y = ((x * r) * w).float().view(B, L, 24, 64, 2)
    t = torch.arange(L, device="cuda")
    f, h, ww = t // (gh * gw), (t // gw) % gh, t % gw
    a = torch.cat([ang[f, :22], ang[h, 22:43], ang[ww, 43:]], -1)  # [L, 64]
    c, s = a.cos().float()[None, :, None], a.sin().float()[None, :, None]
    re = y[..., 0] * c - y[..., 1] * s
    im = y[..., 0] * s + y[..., 1] * c
    want = torch.stack([re, im], -1).reshape(B * L, D)
    e1 = _rel(buf[:, :D], want)
    untouched = torch.equal(buf[:, D:], buf[:, D:])
    return e1 < 5e-3 and untouched, {"rel": e1}


CHECKS = {
    "ln": lambda: chk_ln(),
    "rms": lambda: chk_rms(),
    "gemm_small": lambda: chk_gemm(128, 256, 64, 0, bias=False),
    "gemm_k128": lambda: chk_gemm(128, 256, 128, 0, bias=False),
    "gemm_k512": lambda: chk_gemm(256, 512, 512, 0),
    "gemm_bn128": lambda: chk_gemm(256, 128, 256, 0),
    "gemm_bn192": lambda: chk_gemm(256, 192, 256, 0),
    "gemm_bn64": lambda: chk_gemm(256, 64, 256, 0),
    "gemm_tail": lambda: chk_gemm(1344, 3072, 592, 0),
    "gemm_big": lambda: chk_gemm(4096, 3072, 3072, 0),
    "gemm_gelu": lambda: chk_gemm(512, 1024, 256, 1),
    "gemm_f32": lambda: chk_gemm(512, 512, 256, 2),
    "gemm_resid": lambda: chk_gemm(1344, 3072, 512, 3),
    "fmha_1tile": lambda: chk_fmha(1, 1, 256, 128),
    "fmha_2tile": lambda: chk_fmha(1, 1, 256, 256),
    "fmha_tail": lambda: chk_fmha(1, 2, 200, 672),
    "fmha_multi": lambda: chk_fmha(2, 3, 672, 672),
    "fmha_cross": lambda: chk_fmha(2, 24, 1344, 512),
    "fmha_long": lambda: chk_fmha(1, 2, 1024, 11648),
}

if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--one":
        ok, info = CHECKS[sys.argv[2]]()
        print("RESULT " + json.dumps({"check": sys.argv[2], "ok": bool(ok), "info": info}))
        sys.exit(0)
    names = sys.argv[1:] or list(CHECKS)
    results = []
    for n in names:
        try:
            r = subprocess.run([sys.executable, __file__, "--one", n], capture_output=True, text=True, timeout=60)
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")]
            if line:
                res = json.loads(line[-1][7:])
            else:
                res = {"check": n, "ok": False, "info": {"rc": r.returncode, "stderr": r.stderr[-1500:]}}
        except subprocess.TimeoutExpired:
            res = {"check": n, "ok": False, "info": "TIMEOUT (deadlock?)"}
        print(json.dumps(res), flush=True)
        results.append(res)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w") as f:
        json.dump(results, f, indent=1)
    print("SUMMARY ok=%d fail=%d" % (sum(r["ok"] for r in results), sum(not r["ok"] for r in results)))
