"""GPU: the tcgen05 fp16 x 3 split building blocks (dagnn_b200/csrc/tc.cuh) against fp64 matmul.
fp32-grade accuracy is the point: a single fp16 / TF32 pass would miss the bound by ~100x."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("R,N,K", [(48, 128, 256), (64, 256, 256), (57, 96, 208), (16, 16, 16), (48, 64, 64), (33, 240, 144)])
@pytest.mark.parametrize("scale", [1.0, 0.05])
def test_tc_selftest_ts_matches_fp64(built_lib, R, N, K, scale):
    """The cluster sweep's arrangement: weights resident in TMEM as the A operand ([hi ; lo] stacked on the lanes), the
    rows of the level as the shared-memory B operand. C[N,R] = X[N,K] W[R,K]^T at fp32-grade accuracy."""
    from dagnn_b200 import _lib
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(R * 13 + N)
    W = (torch.randn(R, K, generator=g) * scale).to(dev)
    X = torch.randn(N, K, generator=g).to(dev)
    C = torch.full((N, R), float("nan"), device=dev)
    _lib.check(built_lib.dagnn_tc_selftest_ts(W.data_ptr(), X.data_ptr(), C.data_ptr(), R, N, K,
                                              torch.cuda.current_stream().cuda_stream), "dagnn_tc_selftest_ts")
    torch.cuda.synchronize()
    ref = X.double() @ W.double().t()
    err = (C.double() - ref).abs().max().item()
    scale_ref = ref.abs().max().item()
    assert err <= 1e-5 * scale_ref + 1e-5, "max-abs err %g (scale %g)" % (err, scale_ref)


@pytest.mark.parametrize("M,N,K", [(128, 192, 256), (300, 48, 64), (77, 128, 512), (128, 256, 128), (1000, 16, 64), (200, 96, 640)])
@pytest.mark.parametrize("scale", [1.0, 0.05])
def test_tc_selftest_f16x3_matches_fp64(built_lib, M, N, K, scale):
    """fp16 x 3 split (the level kernel's arithmetic): weights of magnitude ~1/sqrt(H) (scale 0.05) put the lo parts into
    the fp16 subnormal range; the absolute error must stay at the fp32 level."""
    from dagnn_b200 import _lib
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M * 11 + N)
    A = torch.randn(M, K, generator=g).to(dev)
    B = (torch.randn(N, K, generator=g) * scale).to(dev)
    C = torch.full((M, N), float("nan"), device=dev)
    _lib.check(built_lib.dagnn_tc_selftest_f16x3(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K,
                                                 torch.cuda.current_stream().cuda_stream), "dagnn_tc_selftest_f16x3")
    torch.cuda.synchronize()
    ref = (A.double() @ B.double().t())
    err = (C.double() - ref).abs().max().item()
    scale_ref = ref.abs().max().item()
    assert err <= 1e-5 * scale_ref + 1e-5, "max-abs err %g (scale %g)" % (err, scale_ref)
