"""GPU-side diagnostic: run small GEMMs through the C ABI and print where they differ from torch."""
import ctypes as C
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pcseg_b200

lib = pcseg_b200._lib
torch.backends.cuda.matmul.allow_tf32 = False


def run(layout, M, N, K, bn=0, structured=False):
    torch.manual_seed(0)
    if layout == 0:
        A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
        B = (torch.randn(N, K, device="cuda") * 0.5).bfloat16()
        if structured:
            A = (torch.arange(M, device="cuda")[:, None] % 7 + 1).float().expand(M, K).contiguous().bfloat16() * 0 + \
                torch.eye(M, K, device="cuda").bfloat16()
        bias = torch.zeros(N, device="cuda")
        D = torch.full((M, N), -3.0, device="cuda").bfloat16()
        rc = lib.pcseg_gemm_test(0, M, N, K, C.c_void_p(A.data_ptr()), K, C.c_void_p(B.data_ptr()), K, C.c_void_p(D.data_ptr()), N,
                                 C.c_void_p(bias.data_ptr()), bn, None)
        ref = torch.relu(A.float() @ B.float().t())
        got = D.float()
    else:
        A = (torch.randn(K, M, device="cuda") * 0.5).bfloat16()
        B = (torch.randn(K, N, device="cuda") * 0.5).bfloat16()
        D = torch.zeros(M, N, device="cuda")
        rc = lib.pcseg_gemm_test(1, M, N, K, C.c_void_p(A.data_ptr()), M, C.c_void_p(B.data_ptr()), N, C.c_void_p(D.data_ptr()), N,
                                 None, bn, None)
        ref = A.float().t() @ B.float()
        got = D
    if rc != 0:
        print("  rc", rc, lib.pcseg_last_error().decode())
        return
    torch.cuda.synchronize()
    err = (got - ref).abs()
    print(f"layout {layout} M={M} N={N} K={K} bn={bn}: max err {err.max().item():.4g} (ref max {ref.abs().max().item():.4g})")
    if err.max().item() > 0.05 * ref.abs().max().item():
        rb = min(M, 128) // 8 if M >= 8 else 1
        blk = err[: (M // 8) * 8, : (N // 8) * 8].reshape(M // 8, 8, N // 8, 8).amax(dim=(1, 3))
        print("  8x8 block error map (first 16x16 blocks):")
        for r in range(min(16, blk.shape[0])):
            print("   ", " ".join(f"{v:6.2f}" for v in blk[r, :16].tolist()))
        print("  got[0,:8]", got[0, :8].tolist())
        print("  ref[0,:8]", ref[0, :8].tolist())


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "k"):
        run(0, 128, 64, 64)
        run(0, 128, 64, 128)
        run(0, 256, 256, 64)
        run(0, 300, 128, 256)
    if which in ("all", "mn"):
        run(1, 128, 64, 64)
        run(1, 128, 64, 256)
        run(1, 64, 64, 1000)
        run(1, 256, 256, 512)
