"""tcgen05 GEMM kernel vs a plain fp32 torch matmul of the same bf16 inputs (through the C ABI)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    import pcseg_b200
    from pcseg_b200 import _lib as L   # noqa
    return pcseg_b200._lib


def _call(layout, M, N, K, A, lda, B, ldb, D, ldc, bias, bn):
    lib = _lib()
    s = torch.cuda.current_stream().cuda_stream
    rc = lib.pcseg_gemm_test(layout, M, N, K, C.c_void_p(A.data_ptr()), lda, C.c_void_p(B.data_ptr()), ldb,
                             C.c_void_p(D.data_ptr()), ldc, None if bias is None else C.c_void_p(bias.data_ptr()), bn,
                             C.c_void_p(s))
    assert rc == 0, lib.pcseg_last_error().decode()
    torch.cuda.synchronize()


@pytest.mark.parametrize("M,N,K,bn", [
    (128, 64, 64, 0), (300, 64, 64, 0), (1000, 128, 64, 0), (257, 256, 128, 0), (4096, 1024, 1024, 0),
    (700, 512, 64, 64), (700, 512, 576, 128), (33, 256, 256, 256), (20000, 256, 512, 0),
])
def test_gemm_kmajor_bias_relu(M, N, K, bn):
    torch.manual_seed(M + N + K)
    torch.backends.cuda.matmul.allow_tf32 = False
    A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    B = (torch.randn(N, K, device="cuda") * 0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    D = torch.full((M, N), 7.0, device="cuda").bfloat16()
    _call(0, M, N, K, A, K, B, K, D, N, bias, bn)
    ref = torch.relu(A.float() @ B.float().t() + bias)
    err = (D.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 1e-2 * scale + 1e-3, (err, scale)


def test_gemm_kmajor_strided_operands():
    """A with a row pitch larger than K (the [dy3 | dy_seg1] concatenated buffer uses this)."""
    torch.manual_seed(1)
    M, N, K, lda = 500, 64, 128, 576
    Abuf = (torch.randn(M, lda, device="cuda") * 0.5).bfloat16()
    B = (torch.randn(N, K, device="cuda") * 0.5).bfloat16()
    bias = torch.zeros(N, device="cuda")
    D = torch.zeros(M, N, device="cuda").bfloat16()
    A = Abuf[:, 64:64 + K]
    _call(0, M, N, K, A, lda, B, K, D, N, bias, 0)
    ref = torch.relu(A.float() @ B.float().t())
    assert (D.float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item() + 1e-3


@pytest.mark.parametrize("P,Mc,Nc,bn", [
    (64, 128, 64, 0), (1000, 64, 64, 0), (5000, 128, 256, 0), (3000, 512, 64, 0), (2049, 1024, 1024, 0),
    (777, 256, 512, 128), (40000, 128, 128, 0),
])
def test_gemm_wgrad_mn_major(P, Mc, Nc, bn):
    torch.manual_seed(P + Mc + Nc)
    torch.backends.cuda.matmul.allow_tf32 = False
    A = (torch.randn(P, Mc, device="cuda") * 0.5).bfloat16()      # dY  [points, Cout]
    B = (torch.randn(P, Nc, device="cuda") * 0.5).bfloat16()      # act [points, Cin]
    D = torch.zeros(Mc, Nc, device="cuda")
    _call(1, Mc, Nc, P, A, Mc, B, Nc, D, Nc, None, bn)
    ref = A.float().t() @ B.float()
    err = (D - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-3 * scale + 1e-3, (err, scale)


def test_gemm_wgrad_pitched_destination():
    """seg_conv1.weight gradient: destination pitch 1088, source columns 64.. of a 576-wide buffer."""
    torch.manual_seed(3)
    P, Mc, Nc = 1500, 512, 64
    dycat = (torch.randn(P, 576, device="cuda") * 0.5).bfloat16()
    act = (torch.randn(P, Nc, device="cuda") * 0.5).bfloat16()
    D = torch.zeros(Mc, 1088, device="cuda")
    A = dycat[:, 64:]
    _call(1, Mc, Nc, P, A, 576, act, Nc, D, 1088, None, 0)
    ref = A.float().t() @ act.float()
    assert (D[:, :64] - ref).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-3
    assert D[:, 64:].abs().max().item() == 0.0
