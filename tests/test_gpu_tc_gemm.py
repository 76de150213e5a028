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


@pytest.mark.parametrize("S,R,C,spec", [(1, 128, 32, [64]), (3, 1000, 259, [128, 128, 128]), (2, 16384, 260, [128, 128, 128]),
                                        (5, 77, 37, [64, 40]), (8, 4096, 128, [256])])
def test_row_mlp_channel_major_source(pkg, S, R, C, spec):
    """b200pn2_row_mlp_forward_cm reads the (S, C, R) conv layout in place: same numbers as transposing first (the producers
    deliver the same operand values to the same MMAs), odd channel counts and ragged last tiles included."""
    import pointnet2._ext as ext
    g = torch.Generator(device="cuda").manual_seed(S * 100 + C)
    x_cm = torch.randn(S, C, R, device="cuda", generator=g)
    layers, cin = [], C
    for co in spec:
        layers.append((torch.randn(co, cin, device="cuda", generator=g) / cin ** 0.5,
                       torch.rand(co, device="cuda", generator=g) + 0.5, torch.randn(co, device="cuda", generator=g) * 0.1))
        cin = co
    a_cm, a_pm = ext.row_mlp_forward_cm(x_cm, layers, relu_last=True, want_cm=True, want_pm=True)
    ld = (C + 3) // 4 * 4
    rows = ext.transpose_cn(x_cm, ld=ld)
    b_cm, b_pm = ext.row_mlp_forward(rows, layers, relu_last=True, want_cm=True, want_pm=True, channels=C)
    assert torch.equal(a_cm, b_cm) and torch.equal(a_pm, b_pm)
    ref = x_cm.transpose(1, 2).double()
    for w, sc, sh in layers:
        ref = ((ref @ w.double().t()) * sc.double() + sh.double()).clamp_min(0)
    err = (a_pm.double() - ref).abs()
    assert bool((err <= 1e-5 + 1e-5 * ref.abs()).all()), float(err.max())
