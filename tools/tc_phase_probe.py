"""Where does a tiled tcgen05 GEMM CTA spend its time?  Needs the instrumented build:
    CB200_EXTRA_NVCC_FLAGS=-DCB200_TC_PROF python -m coach_b200.build --force
Runs conv2 / conv3 / fc1 shaped forward GEMMs of the Atari network at B = 512 on pre-split planes and prints the cycles
the producer lane 0, the MMA thread and epilogue thread 0 of CTA (0,0,0) spent per phase."""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from coach_b200 import _lib                                      # noqa: E402
from coach_b200.architectures import tiled as tl                 # noqa: E402
from coach_b200.architectures.layers import Conv2d, Dense, Workspace   # noqa: E402

lib = _lib.load()
raw = ctypes.CDLL(_lib.LIB_PATH)
dev = torch.device("cuda")
ws = Workspace(dev)


def read(reset=True):
    out = (ctypes.c_ulonglong * 16)()
    raw.cb200_tc_prof_read(out, int(reset))
    return list(out)


def run(layer, B, npix_in, C, tag):
    g = torch.Generator().manual_seed(0)
    K = int(np.prod(layer.param_shapes[0][1][:-1]))
    N = layer.N
    x = torch.relu(torch.randn(B, npix_in * C, generator=g)).to(dev)
    w = (torch.randn(K, N, generator=g) / np.sqrt(K)).to(dev)
    b = torch.zeros(N, device=dev)
    y = torch.empty(B, layer.out_elems(), device=dev)
    wp = tl.PlaneBuf(K, N, dev).load(lib, w)
    xp = tl.PlaneBuf(npix_in * B, C, dev, npix=npix_in).load(lib, x.view(B, npix_in, C).permute(1, 0, 2))
    ctx = tl.PlaneCtx(x=xp, y=tl.PlaneBuf(layer.out_pixels() * B, N, dev, npix=layer.out_pixels()), w_ptr=wp.ptr,
                      w_stride=wp.stride)
    layer.prepare(lib, ws, B, dev, x, y, w, b, None, None, None, None, need_dx=False, planes=ctx)
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
    chunks = max(1, v[2])
    print("%s: %.1f us/launch, %d chunks/CTA, splits %d" % (tag, e0.elapsed_time(e1) * 1000 / n, chunks // n,
                                                          layer.fwd.desc.splits))
    print("   producer: wait free stage %6.0f   issue copies %6.0f   cycles per chunk" % (v[0] / chunks, v[1] / chunks))
    print("   mma     : wait data       %6.0f   issue mmas   %6.0f   cycles per chunk" % (v[3] / chunks, v[4] / chunks))
    print("   epilogue warps: main loop %8.0f   epilogue %8.0f   cycles per launch" % (v[5] / n, v[6] / n))


B = 512
run(Conv2d((20, 20), 32, 64, 4, 2, "relu"), B, 400, 32, "conv2 fwd")
run(Conv2d((9, 9), 64, 64, 3, 1, "relu"), B, 81, 64, "conv3 fwd")
run(Dense(3136, 512, "relu"), B, 49, 64, "fc1 fwd")
