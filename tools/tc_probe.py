"""Accuracy probe of the tcgen05 3xBF16 GEMM (tools/tc_probe.cu) against fp64; run under gpurun."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, "coach_b200", "lib", "libtcprobe.so"))
lib.tc_probe.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 4 + [ctypes.c_void_p]
lib.tc_probe.restype = ctypes.c_int


def run(M, N, K, nprod, seed=0, relu_like=False):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(K, N, generator=g) / np.sqrt(K)
    if relu_like:
        A = torch.relu(A)
    Ad, Bd = A.cuda(), B.cuda()
    C = torch.full((M, N), float("nan"), device="cuda")
    rc = lib.tc_probe(Ad.data_ptr(), Bd.data_ptr(), C.data_ptr(), M, N, K, nprod, None)
    torch.cuda.synchronize()
    ref64 = A.double() @ B.double()
    ref32 = (Ad @ Bd).cpu().double()          # cuBLAS fp32 (may use TF32? disabled by default for matmul)
    got = C.cpu().double()
    scale = ref64.abs().max().item()
    return {"M": M, "N": N, "K": K, "products": nprod, "rc": rc,
            "max_abs_err_vs_fp64": float((got - ref64).abs().max()), "scale": scale,
            "rel_to_scale": float((got - ref64).abs().max() / scale),
            "fp32_cublas_rel_to_scale": float((ref32 - ref64).abs().max() / scale),
            "nan": bool(torch.isnan(got).any())}


if __name__ == "__main__":
    torch.backends.cuda.matmul.allow_tf32 = False
    for (M, N, K) in [(128, 64, 32), (128, 64, 64), (256, 64, 512), (512, 32, 256), (1024, 64, 576), (512, 128, 3136),
                      (512, 256, 3136), (128, 16, 48)]:
        for nprod in (1, 3, 6):
            print(json.dumps(run(M, N, K, nprod)), flush=True)
    print(json.dumps(run(512, 64, 3136, 6, relu_like=True)))
