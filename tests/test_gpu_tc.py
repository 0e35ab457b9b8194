"""GPU: the tcgen05 3xTF32 building blocks (dagnn_b200/csrc/tc.cuh) against fp64 matmul.
fp32-grade accuracy is the point: single-pass TF32 would miss the bound by ~100x."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(128, 192, 256), (300, 64, 32), (77, 128, 512), (128, 256, 96), (1000, 16, 64)])
def test_tc_selftest_matches_fp64(built_lib, M, N, K):
    from dagnn_b200 import _lib
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M * 7 + N)
    A = torch.randn(M, K, generator=g).to(dev)
    B = torch.randn(N, K, generator=g).to(dev)
    C = torch.full((M, N), float("nan"), device=dev)
    _lib.check(built_lib.dagnn_tc_selftest_f32(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K,
                                               torch.cuda.current_stream().cuda_stream), "dagnn_tc_selftest_f32")
    torch.cuda.synchronize()
    ref = (A.double() @ B.double().t())
    err = (C.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    # the tensor core accumulates in fp32 with round-toward-zero: ~0.5 ulp of bias per accumulated MMA (3 * K/8 of them),
    # i.e. ~6e-6 relative at K = 512 (measured 4.3e-6) - still 100x tighter than single-pass TF32 (5e-4)
    assert err <= 1e-5 * scale + 1e-5, "max-abs err %g (scale %g)" % (err, scale)


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
