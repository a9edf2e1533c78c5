"""Where does a tcgen05 planes-GEMM CTA spend its time?  Needs the instrumented build:
    CB200_EXTRA_NVCC_FLAGS=-DCB200_TC_PROF python -m coach_b200.build --force
Runs the conv2 / conv3 / fc1 shaped forward GEMMs of the Atari network at B = 512 and prints the cycles thread 0 of
CTA (0,0,0) spent per phase of the chunk loop (averaged per chunk)."""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from coach_b200 import _lib                                      # noqa: E402
from coach_b200.architectures.layers import PLANES, Conv2d, Dense, Workspace   # noqa: E402

lib = _lib.load()
raw = ctypes.CDLL(_lib.LIB_PATH)
dev = torch.device("cuda")
ws = Workspace(dev)
names = ["prologue", "wait cp.async", "fence+sync", "mma issue", "wait prev mma", "issue cp.async", "chunks",
         "wait last mma", "epilogue", "total"]


def read(reset=True):
    out = (ctypes.c_ulonglong * 16)()
    raw.cb200_tc_prof_read(out, int(reset))
    return list(out)


def run(layer, B, x_shape, tag):
    g = torch.Generator().manual_seed(0)
    x = torch.relu(torch.randn(*x_shape, generator=g)).to(dev)
    K = int(np.prod([s for _, s in layer.param_shapes][0][:-1]))
    N = layer.N
    w = (torch.randn(K, N, generator=g) / np.sqrt(K)).to(dev)
    b = torch.zeros(N, device=dev)
    y = torch.empty(B, layer.out_elems(), device=dev)
    for t in (x, w):
        PLANES.register(t)
        PLANES.refresh(lib, t)
    PLANES.register(y)
    layer.prepare(lib, ws, B, dev, x, y, w, b, None, None, None, None, need_dx=False, planes=True)
    for _ in range(3):
        layer.forward()
    read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        layer.forward()
    e1.record()
    torch.cuda.synchronize()
    v = read()
    chunks = max(1, v[6])
    print("%s: %.1f us/launch, %d chunks/CTA" % (tag, e0.elapsed_time(e1) * 1000 / n, chunks // n))
    for i, nm in enumerate(names):
        if i in (0, 7, 8, 9):
            print("   %-16s %8.0f cycles per launch" % (nm, v[i] / n))
        elif i != 6:
            print("   %-16s %8.0f cycles per chunk" % (nm, v[i] / chunks))


B = 512
run(Conv2d((20, 20), 32, 64, 4, 2, "relu"), B, (B, 20, 20, 32), "conv2 fwd")
run(Conv2d((9, 9), 64, 64, 3, 1, "relu"), B, (B, 9, 9, 64), "conv3 fwd")
run(Dense(3136, 512, "relu"), B, (B, 3136), "fc1 fwd")
