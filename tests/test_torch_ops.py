"""torch.library registration of the block-level operators (flexam_b200/torch_ops.py):
traceable with fake tensors, CUDA-only (no CPU fallback), and - on the GPU - the same results as flexam_b200.ops."""
import pytest
import torch
from torch._subclasses.fake_tensor import FakeTensorMode

from flexam_b200 import torch_ops

bf16, f32 = torch.bfloat16, torch.float32


def test_every_operator_is_registered():
    for name in torch_ops.OPERATORS:
        op = getattr(torch.ops.flexam_b200, name)
        assert op.default._schema.is_mutable, name            # writes into a caller-provided tensor
        assert str(op.default._schema.returns) == "[]", name  # and returns nothing, like the C ABI


def test_fake_tensor_tracing_of_a_block_slice():
    """A fragment of WanAttentionBlock (LN+modulate -> qkv GEMM -> q|k norm+RoPE -> attention -> gated o-projection)
    traces under FakeTensorMode without touching a device or the shared library."""
    B, L, H, D = 2, 96, 2, 256
    with FakeTensorMode():
        def t(*shape, dtype=bf16):
            return torch.empty(*shape, dtype=dtype, device="cuda")
        x, h = t(B * L, D, dtype=f32), t(B * L, D)
        tab = t(2, 2 * B, 2, D, dtype=f32)
        idx = torch.empty(B * L, dtype=torch.int32, device="cuda")
        torch.ops.flexam_b200.ln_scale_shift(x, h, 1e-6, tab[0, :, 0], tab[0, :, 1], 2 * D, idx)
        qkv, w = t(B * L, 3 * D), t(3 * D, D)
        torch.ops.flexam_b200.gemm(h, w, t(3 * D), qkv, 0)
        torch.ops.flexam_b200.rmsnorm_rope(qkv[:, :2 * D], t(D), 1e-6, t(1024, 64, 2, dtype=f32), [3, 4, 8], 0, L, t(D))
        v5 = qkv.view(B, L, 3, H, 128)
        attn = t(B, L, H, 128)
        torch.ops.flexam_b200.fmha(v5[:, :, 0], v5[:, :, 1], v5[:, :, 2], attn, 128 ** -0.5)
        torch.ops.flexam_b200.gemm(attn.view(B * L, D), t(D, D), t(D), x, 3, t(D, dtype=f32), t(2, D, dtype=f32), idx)
        assert x.shape == (B * L, D) and x.dtype == f32 and attn.dtype == bf16


def test_no_cpu_fallback_behind_the_dispatcher():
    a, w, out = torch.zeros(4, 8, dtype=bf16), torch.zeros(16, 8, dtype=bf16), torch.zeros(4, 16, dtype=bf16)
    with pytest.raises(NotImplementedError):
        torch.ops.flexam_b200.gemm(a, w, None, out, 0)
    with pytest.raises(NotImplementedError):
        torch.ops.flexam_b200.ln_affine(torch.zeros(4, 256), torch.zeros(4, 256, dtype=bf16), 1e-6,
                                        torch.ones(256, dtype=bf16), torch.zeros(256, dtype=bf16))


@pytest.mark.gpu
def test_dispatcher_ops_match_direct_binding():
    from flexam_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(11)
    M, N, K = 300, 512, 256
    a = torch.randn(M, K, device=dev, generator=g).bfloat16()
    w = (torch.randn(N, K, device=dev, generator=g) * K ** -0.5).bfloat16()
    b = torch.randn(N, device=dev, generator=g).bfloat16()
    o1, o2 = torch.empty(M, N, device=dev, dtype=bf16), torch.empty(M, N, device=dev, dtype=bf16)
    ops.gemm(a, w, b, o1, ops.FX_EPI_GELU_BF16)
    torch.ops.flexam_b200.gemm(a, w, b, o2, ops.FX_EPI_GELU_BF16)
    assert torch.equal(o1, o2)
    qkv = torch.randn(2 * 200, 3 * 256, device=dev, generator=g).bfloat16().view(2, 200, 3, 2, 128)
    a1, a2 = torch.empty(2, 200, 2, 128, device=dev, dtype=bf16), torch.empty(2, 200, 2, 128, device=dev, dtype=bf16)
    ops.fmha(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], a1, 128 ** -0.5)
    torch.ops.flexam_b200.fmha(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], a2, 128 ** -0.5)
    assert torch.equal(a1, a2)
    x = torch.randn(64, 256, device=dev, generator=g)
    gam, bet = torch.ones(256, device=dev, dtype=bf16), torch.zeros(256, device=dev, dtype=bf16)
    l1, l2 = torch.empty(64, 256, device=dev, dtype=bf16), torch.empty(64, 256, device=dev, dtype=bf16)
    ops.ln_affine(x, l1, 1e-6, gam, bet)
    torch.ops.flexam_b200.ln_affine(x, l2, 1e-6, gam, bet)
    assert torch.equal(l1, l2)
