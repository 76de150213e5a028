"""GPU: pins the tcgen05 building blocks of the product kernel (descriptor encodings, 128-byte swizzle, TMEM addressing,
commit/mbarrier protocol, split-precision accumulation) through its plainest use: b200pn2_row_mlp_forward with ONE layer
and identity affine is a (rows x K) . (K x N) GEMM, checked against an fp64 reference."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def run(pkg, rows, N, K, seed=0, relu=False):
    import pointnet2._ext as ext
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(1, rows, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    one, zero = torch.ones(N, device="cuda"), torch.zeros(N, device="cuda")
    out_cm, out_pm = ext.row_mlp_forward(A, [(W, one, zero)], relu_last=relu, want_cm=True, want_pm=True)
    torch.cuda.synchronize()
    ref = A[0].double() @ W.double().t()
    if relu:
        ref = ref.clamp_min(0)
    assert torch.equal(out_cm[0].t().contiguous(), out_pm[0])          # both layouts hold the same numbers
    return (out_pm[0].double() - ref).abs().max().item(), ref.abs().max().item()


@pytest.mark.parametrize("rows,N,K", [(128, 128, 32), (1000, 128, 128), (4096, 256, 128), (300, 128, 288), (513, 256, 160),
                                      (777, 97, 128), (256, 3, 256), (40000, 64, 128)])
def test_split_tf32_row_gemm(pkg, rows, N, K):
    err, scale = run(pkg, rows, N, K)
    assert err < 3e-6 * max(scale, 1.0), "split-precision result not fp32-accurate: %g (scale %g)" % (err, scale)


def test_row_gemm_relu_and_unaligned_channels(pkg):
    err, scale = run(pkg, 999, 128, 259, relu=True)   # K not a multiple of 4: scalar gather tail
    assert err < 3e-6 * max(scale, 1.0)
