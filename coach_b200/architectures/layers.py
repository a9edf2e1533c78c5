"""Layers of the learn-step networks as prepared gather-GEMM calls (include/coach_b200.h: cb200_gemm).

Mirrors the layer vocabulary of ``rl_coach/architectures/tensorflow_components/layers.py:108-183`` (Conv2d, Dense) and
the embedders' input rescale (``embedders/embedder.py:103-104``).  Data layout is TensorFlow's: activations NHWC,
conv kernels HWIO (= a row-major [KH*KW*Cin, N] matrix), dense kernels [in, out]; VALID padding
(``tf.layers.conv2d`` default).

Every contraction is ONE C-ABI call whose descriptor -- index tables included -- is built once per (layer, batch size)
and reused every step: forward, weight gradient (A^T * dZ with a fixed-order split reduction over the batch*pixels
axis; the bias gradient rides along as one extra output row when the bias gradient sits right behind the kernel
gradient in the flat buffer), data gradient (dense: dZ * W^T; conv: transposed convolution in gather form, one call
per stride-parity class, with the previous layer's activation derivative fused into the epilogue).
"""
import ctypes

import numpy as np
import torch

from coach_b200 import _lib

ACT = {None: 0, "none": 0, "relu": 1, "tanh": 2}


def _dev_i32(a, device):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(device)


class Workspace(object):
    """One scratch buffer shared by all ops of a network (split-reduction partials, column sums)."""

    def __init__(self, device):
        self.device = device
        self.buf = torch.empty(1 << 20, dtype=torch.float32, device=device)

    def require(self, nfloats):
        if self.buf.numel() < nfloats:
            self.buf = torch.empty(int(nfloats), dtype=torch.float32, device=self.device)

    def ptr(self):
        return self.buf.data_ptr()


class PlaneRegistry(object):
    """fp32 device buffers that are shadowed by bf16 hi / mid / lo planes (cb200_gemm_desc.a_planes / b_planes /
    c_planes).  A GemmOp whose operand lies inside a registered buffer picks the planes up automatically; whoever
    writes a registered buffer outside a GEMM epilogue must refresh the planes (``refresh``)."""

    def __init__(self):
        self.entries = []            # (weakref to base tensor, base_ptr, nbytes, planes [3, stride] bf16)

    def register(self, t):
        import weakref
        assert t.dtype == torch.float32 and t.is_contiguous()
        if t.numel() % 8 or t.data_ptr() % 16:
            return None
        self.entries = [e for e in self.entries if e[0]() is not None]
        for ref, ptr, nbytes, planes in self.entries:
            if ptr == t.data_ptr() and nbytes == t.numel() * 4:
                return planes
        stride = (t.numel() + 7) // 8 * 8
        planes = torch.zeros((3, stride), dtype=torch.bfloat16, device=t.device)
        self.entries.append((weakref.ref(t), t.data_ptr(), t.numel() * 4, planes))
        return planes

    def lookup(self, ptr):
        """(pointer to the plane-0 element shadowing `ptr`, plane stride in elements) or (0, 0)"""
        for ref, base, nbytes, planes in self.entries:
            if ref() is not None and base <= ptr < base + nbytes:
                e = (ptr - base) // 4
                if e % 8 == 0 and planes.data_ptr() % 16 == 0:
                    return planes.data_ptr() + 2 * e, planes.shape[1]
        return 0, 0

    def refresh(self, lib, t):
        """re-derive the planes of a registered buffer from its fp32 content (one launch)"""
        planes = self.register(t)
        n = t.numel()
        _lib.check(lib.cb200_split_planes(t.data_ptr(), n, planes.data_ptr(), planes.shape[1], _lib.current_stream()))
        return planes


PLANES = PlaneRegistry()


def _u8_div(lut, x_is_u8):
    return float(getattr(lut, "u8_div", 0.0)) if (x_is_u8 and lut is not None) else 0.0


class GemmOp(object):
    """A prepared cb200_gemm call.  Tensors referenced by the descriptor are kept alive here."""

    def __init__(self, lib, ws, **fields):
        self.lib = lib
        self.ws = ws
        self.keep = []
        self.desc = _lib.GemmDesc()
        self.splits = int(fields.pop("splits", 1))
        use_planes = bool(fields.pop("use_planes", False))
        for k, v in fields.items():
            if torch.is_tensor(v):
                self.keep.append(v)
                v = v.data_ptr()
            setattr(self.desc, k, v)
        self.desc.splits = self.splits
        # operands / outputs that live in plane-shadowed buffers (PlaneRegistry): hand the planes to the kernel
        # (only for ops of an instance that keeps them current: use_planes)
        d = self.desc
        if use_planes and not d.a_lut and d.a_vec8:
            d.a_planes, d.a_plane_stride = PLANES.lookup(d.a_src or 0)
        if use_planes and d.n % 8 == 0 and d.ldb % 8 == 0:
            d.b_planes, d.b_plane_stride = PLANES.lookup(d.b or 0)
        if use_planes and d.ldc % 8 == 0:
            d.c_planes, d.c_plane_stride = PLANES.lookup(d.c or 0)
        rows = self.desc.a_cols if self.desc.a_transposed else self.desc.a_rows
        if self.desc.a_ones_col:
            rows += 1
        if self.splits > 1:
            ws.require(self.splits * rows * self.desc.n)

    def run(self):
        if self.splits > 1:
            self.desc.workspace = self.ws.ptr()
        _lib.check(self.lib.cb200_gemm(ctypes.byref(self.desc), _lib.current_stream()))


def pick_splits(tiles, reduction, sm=148, min_chunk=128):
    """Split the reduction so that the grid has about two CTAs per SM, never below `min_chunk` per split."""
    # the tensor-core path accumulates in TMEM with truncation: never reduce more than TC_MAX_R terms in one launch,
    # the partial sums are then added in fp32 round-to-nearest by the split-reduce kernel (csrc/nn_gemm_tc.cuh)
    need = (reduction + TC_MAX_R - 1) // TC_MAX_R
    if tiles >= sm:          # one full wave already: the extra reduction pass would cost more than it saves
        return int(max(1, need))
    s = max(1, (2 * sm + tiles - 1) // tiles)
    s = min(s, max(1, reduction // min_chunk))
    return int(max(s, need))


TC_MAX_R = 1024


def _tiles(M, N, fast=True):
    """CTA tiles of the kernel cb200_gemm will pick (csrc/nn.cu)."""
    if M <= 64:
        return ((M + 31) // 32) * ((N + 31) // 32)
    if fast and N % 4 == 0:
        if N <= 32:
            return ((M + 255) // 256) * ((N + 31) // 32)
        if N <= 64:
            return ((M + 127) // 128) * ((N + 63) // 64)
        return ((M + 127) // 128) * ((N + 127) // 128)
    bn = 32 if N <= 32 else 64
    return ((M + 127) // 128) * ((N + bn - 1) // bn)


def _bias_rides_along(dw, db, K, N):
    """True when the bias gradient is stored right behind the kernel gradient (flat ParamStore layout)."""
    bm = 256 if N <= 32 else 128          # row tile of the kernel that will run; only use slack of the last tile
    return (dw is not None and db is not None and N % 4 == 0 and K % 4 == 0 and K + 1 > 64 and K % bm != 0 and
            db.data_ptr() == dw.data_ptr() + K * N * 4)


# =====================================================================================================================
class Dense(object):
    """y = act(x W + b), W [K, N] (layers.py:148-183)."""

    def __init__(self, in_features, out_features, activation=None):
        self.K, self.N = int(in_features), int(out_features)
        self.act = ACT[activation]
        self.param_shapes = [("kernel", (self.K, self.N)), ("bias", (self.N,))]
        self.in_shape = (self.K,)
        self.out_shape = (self.N,)

    def out_elems(self):
        return self.N

    def prepare(self, lib, ws, B, device, x, y, w, b, dw, db, dy, dx, x_is_u8=False, lut=None, need_dx=True,
                prev_act=0, dx_accumulate=False, planes=False):
        """x [B,K], y [B,N], dy [B,N] (gradient wrt the PRE-activation of this layer), dx [B,K] (gradient wrt the
        pre-activation of the previous layer: masked with prev_act' evaluated on x)."""
        K, N = self.K, self.N
        rowoff = _dev_i32(np.arange(B) * K, device)
        coloff = _dev_i32(np.arange(K), device)
        vec = int(K % 4 == 0)
        common = dict(a_rowoff=rowoff, a_coloff=coloff, a_rows=B, a_cols=K, a_vec4=vec, a_src=x,
                      a_lut=lut if x_is_u8 else None, a_u8_div=_u8_div(lut, x_is_u8), a_vec8=int(K % 8 == 0),
                      use_planes=planes)
        self.fwd = GemmOp(lib, ws, a_transposed=0, b=w, ldb=N, n=N, c=y, ldc=N, bias=b, act=self.act,
                          splits=pick_splits(_tiles(B, N, vec), K), **common)
        self.bwd_w = None
        self.db_args = None
        if dw is not None and dy is not None:
            ones = int(bool(vec) and _bias_rides_along(dw, db, K, N))
            self.bwd_w = GemmOp(lib, ws, a_transposed=1, b=dy, ldb=N, n=N, c=dw, ldc=N, a_ones_col=ones,
                                splits=pick_splits(_tiles(K + ones, N, vec), B), **common)
            if not ones:
                self.db_args = (dy, B, N, db)
                ws.require(1024 * N)
        self.bwd_x = None
        if need_dx:
            self.wT = torch.empty((N, K), dtype=torch.float32, device=device)
            self.wT_planes = PLANES.register(self.wT) if planes else None
            self.w = w
            ro = _dev_i32(np.arange(B) * N, device)
            co = _dev_i32(np.arange(N), device)
            self.bwd_x = GemmOp(lib, ws, a_src=dy, a_rowoff=ro, a_coloff=co, a_rows=B, a_cols=N, a_transposed=0,
                                a_vec4=int(N % 4 == 0), a_vec8=int(N % 8 == 0), b=self.wT, ldb=K, n=K, c=dx, ldc=K,
                                use_planes=planes,
                                mask_y=x if prev_act else None, mask_act=prev_act,
                                accumulate=int(bool(dx_accumulate)),
                                splits=pick_splits(_tiles(B, K, N % 4 == 0), N))
        self.lib, self.ws = lib, ws

    def forward(self):
        self.fwd.run()

    def backward(self, weights=True):
        st = _lib.current_stream()
        if weights:
            self.bwd_w.run()
            if self.db_args is not None:
                dy, B, N, db = self.db_args
                _lib.check(self.lib.cb200_colsum(dy.data_ptr(), B, N, db.data_ptr(), self.ws.ptr(), st))
        if self.bwd_x is not None:
            pl = self.wT_planes
            _lib.check(self.lib.cb200_transpose(self.w.data_ptr(), self.K, self.N, self.wT.data_ptr(),
                                                pl.data_ptr() if pl is not None else None,
                                                pl.shape[1] if pl is not None else 0, st))
            self.bwd_x.run()


# =====================================================================================================================
class Conv2d(object):
    """NHWC VALID convolution + bias + activation (layers.py:108-146: tf.layers.conv2d(filters, kernel, strides))."""

    def __init__(self, in_hw, in_channels, num_filters, kernel_size, strides, activation="relu"):
        self.H, self.W = int(in_hw[0]), int(in_hw[1])
        self.C, self.N = int(in_channels), int(num_filters)
        self.KH = self.KW = int(kernel_size)
        self.S = int(strides)
        self.OH = (self.H - self.KH) // self.S + 1
        self.OW = (self.W - self.KW) // self.S + 1
        self.K = self.KH * self.KW * self.C
        self.act = ACT[activation]
        self.param_shapes = [("kernel", (self.KH, self.KW, self.C, self.N)), ("bias", (self.N,))]
        self.in_shape = (self.H, self.W, self.C)
        self.out_shape = (self.OH, self.OW, self.N)

    def out_elems(self):
        return self.OH * self.OW * self.N

    def prepare(self, lib, ws, B, device, x, y, w, b, dw, db, dy, dx, x_is_u8=False, lut=None, need_dx=True,
                prev_act=0, dx_accumulate=False, planes=False):
        assert not dx_accumulate, "accumulating data gradients is only wired for Dense layers"
        H, W, C, N, KH, KW, S, OH, OW, K = self.H, self.W, self.C, self.N, self.KH, self.KW, self.S, self.OH, \
            self.OW, self.K
        M = B * OH * OW
        bb, oy, ox = np.meshgrid(np.arange(B), np.arange(OH), np.arange(OW), indexing="ij")
        rowoff = (((bb * H + oy * S) * W + ox * S) * C).reshape(-1)
        ky, kx, cc = np.meshgrid(np.arange(KH), np.arange(KW), np.arange(C), indexing="ij")
        coloff = ((ky * W + kx) * C + cc).reshape(-1)
        assert rowoff.max() + coloff.max() < 2 ** 31
        vec = int(C % 4 == 0)          # (kx, c) runs are contiguous: groups of 4 channels never straddle a pixel
        common = dict(a_rowoff=_dev_i32(rowoff, device), a_coloff=_dev_i32(coloff, device), a_rows=M, a_cols=K,
                      a_src=x, a_lut=lut if x_is_u8 else None, a_vec4=vec, a_u8_div=_u8_div(lut, x_is_u8),
                      a_vec8=int(C % 8 == 0), use_planes=planes)
        self.fwd = GemmOp(lib, ws, a_transposed=0, b=w, ldb=N, n=N, c=y, ldc=N, bias=b, act=self.act,
                          splits=pick_splits(_tiles(M, N, vec), K), **common)
        self.bwd_w = None
        self.db_args = None
        if dw is not None and dy is not None:
            ones = int(bool(vec) and _bias_rides_along(dw, db, K, N))
            self.bwd_w = GemmOp(lib, ws, a_transposed=1, b=dy, ldb=N, n=N, c=dw, ldc=N, a_ones_col=ones,
                                splits=pick_splits(_tiles(K + ones, N, vec), M, min_chunk=512), **common)
            if not ones:
                self.db_args = (dy, M, N, db)
                ws.require(1024 * N)
        self.lib, self.ws = lib, ws
        self.w = w
        self.classes = []
        if not need_dx:
            return
        # transposed convolution, gather form, one GEMM per stride-parity class of input pixels
        w_index = np.arange(KH * KW * C * N).reshape(KH, KW, C, N)
        for py in range(S):
            for px in range(S):
                IH = (H - py + S - 1) // S
                IW = (W - px + S - 1) // S
                TA = (KH - py + S - 1) // S
                TB = (KW - px + S - 1) // S
                if IH <= 0 or IW <= 0 or TA <= 0 or TB <= 0:
                    continue
                b_, i_, j_ = np.meshgrid(np.arange(B), np.arange(IH), np.arange(IW), indexing="ij")
                ro = (((b_ * OH + i_) * OW + j_) * N).reshape(-1)
                rinfo = ((i_ << 16) | j_).reshape(-1)
                rowmap = ((b_ * H + (S * i_ + py)) * W + (S * j_ + px)).reshape(-1)
                a_, t_, n_ = np.meshgrid(np.arange(TA), np.arange(TB), np.arange(N), indexing="ij")
                co = (-(a_ * OW + t_) * N + n_).reshape(-1)
                cinfo = ((a_ << 16) | t_).reshape(-1)
                # B_class[(a, t, n), c] = W[S*a + py, S*t + px, c, n]
                perm = w_index[S * a_ + py, S * t_ + px, :, n_]          # [TA, TB, N, C]
                perm = perm.reshape(-1)
                wt = torch.empty((TA * TB * N, C), dtype=torch.float32, device=device)
                wt_planes = PLANES.register(wt) if planes else None
                op = GemmOp(lib, ws, a_src=dy, a_rowoff=_dev_i32(ro, device), a_coloff=_dev_i32(co, device),
                            a_rowinfo=_dev_i32(rinfo, device), a_colinfo=_dev_i32(cinfo, device), a_oh=OH, a_ow=OW,
                            a_rows=B * IH * IW, a_cols=TA * TB * N, a_transposed=0, a_vec4=int(N % 4 == 0),
                            a_vec8=int(N % 8 == 0), use_planes=planes,
                            b=wt, ldb=C, n=C, c=dx, ldc=C,
                            mask_y=x if prev_act else None, mask_act=prev_act, c_rowmap=_dev_i32(rowmap, device),
                            splits=pick_splits(_tiles(B * IH * IW, C, N % 4 == 0), TA * TB * N))
                self.classes.append((op, wt, _dev_i32(perm, device), wt_planes))

    def forward(self):
        self.fwd.run()

    def backward(self, weights=True):
        st = _lib.current_stream()
        if weights:
            self.bwd_w.run()
            if self.db_args is not None:
                dy, M, N, db = self.db_args
                _lib.check(self.lib.cb200_colsum(dy.data_ptr(), M, N, db.data_ptr(), self.ws.ptr(), st))
        for op, wt, perm, pl in self.classes:
            _lib.check(self.lib.cb200_permute_f32(self.w.data_ptr(), perm.data_ptr(), perm.numel(), wt.data_ptr(),
                                                  pl.data_ptr() if pl is not None else None,
                                                  pl.shape[1] if pl is not None else 0, st))
            op.run()
