"""dnn.py / sgan.py classifier forward on the GPU.

Host side: describe the Keras graph (``NetSpec``), fold BatchNorm, permute the first Dense
kernel from Keras' Flatten order to the device's [branch][h][w][c] order, convert it to bf16,
build Pillow's BICUBIC coefficient tables, and hand everything to libradarml
(rml_net_*).  ``GpuNetClassifier`` then mirrors Keras' ``model.predict([XZ, YZ, XY])``
(dnn.py:371-381 input convention) and adds ``predict_cubes`` (cubes -> label).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from ._lib import check

RADAR_MAX = 255.0          # common.py:31
ACT = {"none": 0, "relu": 1, "lrelu": 2}


@dataclass
class Conv:
    w: np.ndarray              # (3,3,cin,cout) Keras HWIO
    b: np.ndarray
    act: str
    bn: tuple | None = None    # (gamma, beta, moving_mean, moving_var)


@dataclass
class Dense:
    w: np.ndarray              # (in, out)
    b: np.ndarray
    act: str
    bn: tuple | None = None


@dataclass
class NetSpec:
    """Weights of dnn.py:55-91 ('dnn') or sgan.py:157-217 ('sgan_c' / 'sgan_d')."""
    kind: str
    R: int                     # dnn.py:33 RESCALE = 80, sgan.py:39 RESCALE = 128
    n_classes: int
    branches: list = field(default_factory=list)   # 3 towers (xz, yz, xy) of Conv
    dense: list = field(default_factory=list)       # 3 Dense
    bn_eps: float = 1e-3       # keras.layers.BatchNormalization default epsilon
    alpha: float = 0.2         # sgan.py:141 LeakyReLU(alpha=0.2)


def spec_from(obj) -> NetSpec:
    """Duck-typed copy (lets tests hand over oracle.nets.NetParams without importing it here)."""
    return NetSpec(kind=obj.kind, R=obj.R, n_classes=obj.n_classes,
                   branches=[[Conv(l.w, l.b, l.act, l.bn) for l in br] for br in obj.branches],
                   dense=[Dense(d.w, d.b, d.act, d.bn) for d in obj.dense],
                   bn_eps=obj.bn_eps, alpha=obj.alpha)


# --------------------------------------------------------------------------- host preparation
def fold_bn(w, b, bn, eps):
    """Inference BatchNorm after a linear layer: y = g (x - m)/sqrt(v + eps) + beta."""
    w = np.asarray(w, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if bn is None:
        return w, b
    g, beta, m, v = (np.asarray(t, dtype=np.float64) for t in bn)
    s = g / np.sqrt(v + eps)
    return w * s, (b - m) * s + beta


def to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """float -> bfloat16 bit patterns, round-to-nearest-even."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return (((u + 0x7FFF + ((u >> 16) & 1)) >> 16) & 0xFFFF).astype(np.uint16)


def _bicubic(x, a=-0.5):
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_bicubic_tables(in_size: int, out_size: int):
    """Pillow Resample.c precompute_coeffs (BICUBIC, support 2, antialiased when shrinking):
    what Image.resize at dnn.py:243-245 uses.  Returns (K [out][ksize] f64, bounds [out][2])."""
    scale = in_size / out_size
    fscale = max(scale, 1.0)
    sup = 2.0 * fscale
    ksize = int(math.ceil(sup)) * 2 + 1
    K = np.zeros((out_size, ksize), dtype=np.float64)
    B = np.zeros((out_size, 2), dtype=np.int32)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / fscale
        xmin = max(int(center - sup + 0.5), 0)
        xmax = min(int(center + sup + 0.5), in_size) - xmin
        ww = 0.0
        for x in range(xmax):
            w = _bicubic((x + xmin - center + 0.5) * ss)
            K[xx, x] = w
            ww += w
        if ww != 0.0:
            K[xx, :xmax] /= ww
        B[xx] = (xmin, xmax)
    return K, B


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class GpuNetClassifier:
    """The dnn.py / sgan.py classifier resident on one GPU."""

    def __init__(self, spec, engine=None, device: int = 0, chunk: int = 256):
        from .engine import Engine
        self.spec = spec if isinstance(spec, NetSpec) else spec_from(spec)
        self.engine = engine if engine is not None else Engine(device)
        self.chunk = chunk
        self._ws = None
        self._load()

    # ------------------------------------------------------------------ load
    def _load(self):
        sp, eng = self.spec, self.engine
        lib, ctx = eng.lib, eng.ctx
        head = 1 if sp.kind == "sgan_d" else 0
        check(ctx, lib.rml_net_begin(ctx, sp.R, sp.n_classes, head, float(sp.alpha)))
        sx, sy, sz = eng.dims
        shapes = [(sx, sz), (sy, sz), (sx, sy)]          # xz, yz, xy (common.py:40 order)
        keep = []
        for br, (h, w) in enumerate(shapes):
            kh, bh = pil_bicubic_tables(w, sp.R)
            kv, bv = pil_bicubic_tables(h, sp.R)
            keep += [kh, bh, kv, bv]
            check(ctx, lib.rml_net_set_resize_tables(ctx, br, kh.shape[1], _p(kh), _p(bh),
                                                     kv.shape[1], _p(kv), _p(bv)))
        n_layers = len(sp.branches[0])
        for layer in range(n_layers):
            for br in range(3):
                cv = sp.branches[br][layer]
                w, b = fold_bn(cv.w, cv.b, cv.bn, sp.bn_eps)
                w32 = np.ascontiguousarray(w, dtype=np.float32)
                b32 = np.ascontiguousarray(b, dtype=np.float32)
                check(ctx, lib.rml_net_add_conv(ctx, layer, br, w32.shape[2], w32.shape[3],
                                                ACT[cv.act], _p(w32), _p(b32)))
        hw = sp.R
        for _ in range(n_layers):
            hw = (hw + 1) // 2
        cl = sp.branches[0][-1].w.shape[3]
        d1, d2, d3 = sp.dense
        K = 3 * hw * hw * cl
        assert d1.w.shape == (K, 64), (d1.w.shape, K)
        w1, b1 = fold_bn(d1.w, d1.b, d1.bn, sp.bn_eps)
        # Keras Flatten index (h, w, branch*cl + c)  ->  device index [branch][h][w][c]
        w1 = w1.reshape(hw, hw, 3, cl, 64).transpose(2, 0, 1, 3, 4).reshape(K, 64)
        w1t = to_bf16_bits(np.ascontiguousarray(w1.T))                       # [64][K]
        w2, b2 = fold_bn(d2.w, d2.b, d2.bn, sp.bn_eps)
        w3, b3 = fold_bn(d3.w, d3.b, d3.bn, sp.bn_eps)
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)  # noqa: E731
        b1, w2, b2, w3, b3 = f32(b1), f32(w2), f32(b2), f32(w3), f32(b3)
        check(ctx, lib.rml_net_set_dense(ctx, K, _p(w1t), _p(b1), ACT[d1.act], _p(w2), _p(b2),
                                         ACT[d2.act], _p(w3), _p(b3)))
        check(ctx, lib.rml_net_finish(ctx))
        self.K = K
        self.uses_igemm = bool(lib.rml_net_uses_igemm(ctx))

    def _workspace(self):
        import torch
        need = int(self.engine.lib.rml_net_workspace_bytes(self.engine.ctx, self.chunk))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty((need,), device=self.engine.device, dtype=torch.uint8)
        return self._ws

    # ------------------------------------------------------------------ device entry points
    def preprocess_device(self, xz, yz, xy):
        """dnn.py:202-205 + 240-254 on device projections -> images [B,3,R,R] float32."""
        import torch
        eng = self.engine
        # (p - 127.5) / 127.5 in float32 (dnn.py:203), fused into the concat kernel
        eng.set_affine(RADAR_MAX / 2., RADAR_MAX / 2., True)
        try:
            feats = eng.process_samples(xz, yz, xy, scale=True)
        finally:
            eng.set_affine(0.0, RADAR_MAX, True)
        B = feats.shape[0]
        images = torch.empty((B, 3, self.spec.R, self.spec.R), device=eng.device, dtype=torch.float32)
        if B:
            check(eng.ctx, eng.lib.rml_net_resize(eng.ctx, C.c_void_p(feats.data_ptr()), B,
                                                  C.c_void_p(images.data_ptr()), eng._stream()))
        return images

    def forward_images(self, images, want_logits=False, want_tower=False):
        """images [B,3,R,R] float32 CUDA -> (proba [B,C], label [B][, logits][, tower bf16 bits
        [B,K] int16 in [branch][h][w][c] order])."""
        import torch
        eng = self.engine
        B = images.shape[0]
        Cn = self.spec.n_classes
        proba = torch.empty((B, Cn), device=eng.device, dtype=torch.float32)
        label = torch.empty((B,), device=eng.device, dtype=torch.int32)
        logits = torch.empty((B, Cn), device=eng.device, dtype=torch.float32) if want_logits else None
        tower = torch.empty((B, self.K), device=eng.device, dtype=torch.int16) if want_tower else None
        if B:
            ws = self._workspace()
            images = images.contiguous()
            check(eng.ctx, eng.lib.rml_net_forward_images(
                eng.ctx, C.c_void_p(images.data_ptr()), B, C.c_void_p(ws.data_ptr()), ws.numel(),
                C.c_void_p(proba.data_ptr()), C.c_void_p(logits.data_ptr()) if want_logits else None,
                C.c_void_p(label.data_ptr()),
                C.c_void_p(tower.data_ptr()) if want_tower else None, eng._stream()))
        res = (proba, label)
        if want_logits:
            res += (logits,)
        if want_tower:
            res += (tower,)
        return res

    def predict_cubes(self, cubes, mode="max", ijk=None):
        """cubes [B,22,31,176] CUDA (float32 or uint8) -> (proba [B,C] f32, label [B] i32):
        K1 -> K3 -> K4 -> K5."""
        import torch
        eng = self.engine
        cube_u8 = eng._check_cubes(cubes)
        B = cubes.shape[0]
        Cn = self.spec.n_classes
        proba = torch.empty((B, Cn), device=eng.device, dtype=torch.float32)
        label = torch.empty((B,), device=eng.device, dtype=torch.int32)
        if B == 0:
            return proba, label
        md = 0 if mode in ("max", 0) else 1
        if md == 1:
            if ijk is None:
                raise ValueError("slice mode needs ijk [B,3] int32")
            ijk = ijk.to(device=eng.device, dtype=torch.int32).contiguous()
        ws = self._workspace()
        fn = eng.lib.rml_net_predict_u8 if cube_u8 else eng.lib.rml_net_predict
        check(eng.ctx, fn(
            eng.ctx, C.c_void_p(cubes.data_ptr()), B, md,
            C.c_void_p(ijk.data_ptr()) if ijk is not None else None, C.c_void_p(ws.data_ptr()),
            ws.numel(), C.c_void_p(proba.data_ptr()), C.c_void_p(label.data_ptr()), eng._stream()))
        return proba, label

    # ------------------------------------------------------------------ Keras-like host API
    def predict(self, inputs):
        """Keras ``model.predict([XZ, YZ, XY])`` (dnn.py:371-381 input convention): three
        (n, R, R) or (n, R, R, 1) float32 arrays -> (n, C) float32 class probabilities
        (c_model / dnn) or (n, 1) real/fake probability (sgan d_model)."""
        import torch
        eng = self.engine
        XZ, YZ, XY = (np.asarray(a, dtype=np.float32).reshape(len(a), self.spec.R, self.spec.R)
                      for a in inputs)
        images = torch.from_numpy(np.ascontiguousarray(np.stack([XZ, YZ, XY], axis=1))).to(eng.device)
        proba, _ = self.forward_images(images)
        out = proba.cpu().numpy()
        return out[:, :1] if self.spec.kind == "sgan_d" else out

    def preprocess(self, samples):
        """dnn.py:202-254 for a list of (xz, yz, xy) in [0, 255]: -> (n, R, R, 3) float32 with
        channels XZ, YZ, XY (the array the reference splits into model inputs)."""
        import torch
        eng = self.engine
        xz, yz, xy = (torch.from_numpy(np.ascontiguousarray(
            np.stack([np.asarray(t[i], dtype=np.float32) for t in samples]))).to(eng.device)
            for i in range(3))
        return self.preprocess_device(xz, yz, xy).permute(0, 2, 3, 1).contiguous().cpu().numpy()
