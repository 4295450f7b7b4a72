"""How does tcgen05.mma (kind::f16) accumulate in fp32?  One-minute GPU probe through the generic kernel (dfb_dbg_gemm).

Operands exactly representable in the operand format (so every product is exact in fp32 and the only error is the
accumulation), all positive (the accumulator grows monotonically: a truncating accumulate shows up as a NEGATIVE relative
bias that grows linearly with the number of MMAs, a round-to-nearest one as an unbiased error that grows like its square
root).  Printed per K: mean and max relative error of C against the exact (float64) sum, in units of 2^-24, for one
product per k-step (bf16 operands) and for the three-product f16x3 form.
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gemm(A, B, nprod):
    from deepfit_b200 import _lib
    lib = _lib.load()
    M, K = A.shape
    N = B.shape[0]
    out = torch.full((M, N), float("nan"), device="cuda")
    out_t = torch.full((M, N), float("nan"), device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.dfb_dbg_gemm(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), None, M, N, K, 0, nprod, 0, 128,
                                C.c_void_p(out.data_ptr()), C.c_void_p(out_t.data_ptr()), st))
    torch.cuda.synchronize()
    return out


def main():
    g = torch.Generator(device="cuda").manual_seed(1)
    for signed in (False, True):
        for K in (64, 128, 256, 512, 1024, 2048):
            M, N = 256, 128
            A = (torch.rand(M, K, device="cuda", generator=g) + 0.5) / 32      # sums stay far inside the kernels' FP16 range check
            B = (torch.rand(N, K, device="cuda", generator=g) + 0.5) / 32
            if signed:
                A = A * (torch.randint(0, 2, (M, K), device="cuda", generator=g) * 2 - 1)
            line = "%s K=%4d (%3d k-steps):" % ("signed  " if signed else "positive", K, K // 16)
            for nprod, name, cast in ((1, "bf16 x1", torch.bfloat16), (4, "f16x3 one accumulator", None), (4, "f16x3 corrections apart", None)):
                from deepfit_b200 import _lib
                _lib.load().dfb_dbg_set(11, 0 if name.endswith("one accumulator") else 1)
                a, b = (A.to(cast).float(), B.to(cast).float()) if cast is not None else (A, B)
                out = gemm(a.contiguous(), b.contiguous(), nprod).double()
                ref = a.double() @ b.double().t()
                scale = (a.abs().double() @ b.abs().double().t())
                rel = (out - ref) / scale * 2 ** 24
                line += "  %s: mean %+7.2f max|.| %6.2f (x 2^-24 of sum|products|)" % (name, rel.mean().item(), rel.abs().max().item())
            print(line, flush=True)


if __name__ == "__main__":
    main()
