"""One launch of each block-level GEMM shape (for ncu sweeps): qkv, o (resid), ffn1 (gelu), ffn2 (resid)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from flexam_b200 import ops
dev = torch.device("cuda:0")
M = 2 * 11648
g = torch.Generator(device=dev).manual_seed(0)
for name, N, K, epi in (("qkv", 9216, 3072, 0), ("o_resid", 3072, 3072, 3), ("ffn1_gelu", 14336, 3072, 1), ("ffn2_resid", 3072, 14336, 3)):
    a = (torch.randn(M, K, device=dev, generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev, generator=g) * K ** -0.5).bfloat16()
    b = torch.randn(N, device=dev, generator=g).bfloat16()
    out = torch.zeros(M, N, device=dev, dtype=torch.float32 if epi == 3 else torch.bfloat16)
    torch.cuda.synchronize()
    ops.gemm(a, w, b, out, epi)
    torch.cuda.synchronize()
