"""TEST INFRASTRUCTURE — CPU restatement of the dnn.py / sgan.py classifier forward passes.

TensorFlow/Keras (requirements.txt:64, 28) is not installable here, so the networks are
restated layer by layer in float64 numpy/torch following the reference definitions:

  preprocessing   dnn.py:202-205, 240-254 / sgan.py:638-641, 676-690
                  (p - 127.5)/127.5 in float32, then PIL ``Image.resize((R,R), BICUBIC)`` on a
                  mode-'F' image.  ``pil_bicubic_resize`` restates Pillow's Resample.c
                  (precompute_coeffs + two passes, double accumulation, float32 store) and is
                  held bit-exact to Pillow itself in tests/test_nets_oracle.py.
  dnn classifier  dnn.py:45-52, 55-91   3 x [Conv 64 relu, Conv 32 relu] -> concat -> Flatten
                  -> Dense 64 relu -> Dense 64 relu -> Dense C softmax (Dropout = identity)
  sgan c/d model  sgan.py:132-154, 157-217   3 x [Conv 128/64/32 + BN + LeakyReLU(0.2)] ->
                  concat -> Flatten -> 2 x [Dense 64 + BN + LeakyReLU] -> Dense C ->
                  softmax (c_model) or Z/(Z+1), Z = sum exp(logit) (d_model, sgan.py:125-129)

Keras conventions restated: NHWC tensors, Conv2D kernels (kh,kw,cin,cout), Dense kernels
(in,out), 'same' padding with stride 2 pads (0 before, 1 after) on even sizes, Flatten is
row-major over (H, W, 96) with channels ordered xz|yz|xy, BatchNormalization inference uses
the moving statistics with epsilon 1e-3.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

RADAR_MAX = 255.0


# --------------------------------------------------------------------------- PIL bicubic
def _bicubic(x, a=-0.5):
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_coeffs(in_size: int, out_size: int, support: float = 2.0):
    """Pillow src/libImaging/Resample.c precompute_coeffs for the BICUBIC filter.
    Returns (K [out, ksize] float64, bounds [out, 2] = (first input index, tap count))."""
    scale = in_size / out_size
    fscale = max(scale, 1.0)
    sup = support * fscale
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


def pil_bicubic_resize(p: np.ndarray, R: int) -> np.ndarray:
    """Image.fromarray(p).resize((R, R), BICUBIC) for a float32 2-D array: horizontal pass
    then vertical pass, each accumulating in double and storing float32."""
    h, w = p.shape
    Kh, Bh = pil_coeffs(w, R)
    Kv, Bv = pil_coeffs(h, R)
    tmp = np.zeros((h, R), dtype=np.float32)
    for xx in range(R):
        x0, n = Bh[xx]
        acc = np.zeros(h, dtype=np.float64)
        for x in range(n):
            acc += p[:, x0 + x].astype(np.float64) * Kh[xx, x]
        tmp[:, xx] = acc.astype(np.float32)
    out = np.zeros((R, R), dtype=np.float32)
    for yy in range(R):
        y0, n = Bv[yy]
        acc = np.zeros(R, dtype=np.float64)
        for y in range(n):
            acc += tmp[y0 + y, :].astype(np.float64) * Kv[yy, y]
        out[yy] = acc.astype(np.float32)
    return out


def preprocess(samples, R: int) -> np.ndarray:
    """dnn.py:202-205 + 240-254: [(xz,yz,xy)] in [0,255] -> (n, R, R, 3) float32, channels
    XZ, YZ, XY."""
    out = np.zeros((len(samples), R, R, 3), dtype=np.float32)
    for s, t in enumerate(samples):
        for c, p in enumerate(t):
            q = (np.asarray(p, dtype=np.float32) - RADAR_MAX / 2.) / (RADAR_MAX / 2.)
            out[s, :, :, c] = pil_bicubic_resize(q.astype(np.float32), R)
    return out


# --------------------------------------------------------------------------- parameters
@dataclass
class ConvLayer:
    w: np.ndarray                 # (3, 3, cin, cout) float32, Keras HWIO
    b: np.ndarray                 # (cout,)
    act: str                      # 'relu' | 'lrelu' | 'none'
    bn: tuple | None = None       # (gamma, beta, moving_mean, moving_var)


@dataclass
class DenseLayer:
    w: np.ndarray                 # (in, out) float32
    b: np.ndarray
    act: str
    bn: tuple | None = None


@dataclass
class NetParams:
    kind: str                     # 'dnn' | 'sgan_c' | 'sgan_d'
    R: int
    n_classes: int
    branches: list = field(default_factory=list)   # 3 lists of ConvLayer (xz, yz, xy)
    dense: list = field(default_factory=list)       # DenseLayer x 3
    bn_eps: float = 1e-3          # Keras BatchNormalization default
    alpha: float = 0.2            # LeakyReLU slope, sgan.py:141


def _glorot(rng, shape, fan_in, fan_out):
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def _bn(rng, n):
    return (rng.uniform(0.8, 1.2, n).astype(np.float32), rng.normal(0, 0.1, n).astype(np.float32),
            rng.normal(0, 0.2, n).astype(np.float32), rng.uniform(0.5, 1.5, n).astype(np.float32))


def random_dnn(seed=0, n_classes=3) -> NetParams:
    """dnn.py:55-91 with Keras default initialisers (glorot_uniform kernels; biases made
    non-zero so the bias path is exercised)."""
    rng = np.random.default_rng(seed)
    net = NetParams("dnn", 80, n_classes)
    for _ in range(3):
        net.branches.append([
            ConvLayer(_glorot(rng, (3, 3, 1, 64), 9, 9 * 64), rng.normal(0, 0.05, 64).astype(np.float32), "relu"),
            ConvLayer(_glorot(rng, (3, 3, 64, 32), 9 * 64, 9 * 32), rng.normal(0, 0.05, 32).astype(np.float32), "relu")])
    K = 20 * 20 * 96
    net.dense = [DenseLayer(_glorot(rng, (K, 64), K, 64), rng.normal(0, 0.05, 64).astype(np.float32), "relu"),
                 DenseLayer(_glorot(rng, (64, 64), 64, 64), rng.normal(0, 0.05, 64).astype(np.float32), "relu"),
                 DenseLayer(_glorot(rng, (64, n_classes), 64, n_classes), rng.normal(0, 0.05, n_classes).astype(np.float32), "none")]
    return net


def random_sgan(seed=0, n_classes=3, kind="sgan_c", weight_sd=0.02) -> NetParams:
    """sgan.py:157-217; kernels RandomNormal(0, 0.02) (sgan.py:171), BN with non-trivial
    moving statistics as after training."""
    rng = np.random.default_rng(seed)
    net = NetParams(kind, 128, n_classes)
    for _ in range(3):
        layers, cin = [], 1
        for cout in (128, 64, 32):
            layers.append(ConvLayer(rng.normal(0, weight_sd, (3, 3, cin, cout)).astype(np.float32),
                                    rng.normal(0, 0.02, cout).astype(np.float32), "lrelu", _bn(rng, cout)))
            cin = cout
        net.branches.append(layers)
    K = 16 * 16 * 96
    net.dense = [DenseLayer(rng.normal(0, weight_sd, (K, 64)).astype(np.float32), rng.normal(0, 0.02, 64).astype(np.float32), "lrelu", _bn(rng, 64)),
                 DenseLayer(rng.normal(0, weight_sd * 5, (64, 64)).astype(np.float32), rng.normal(0, 0.02, 64).astype(np.float32), "lrelu", _bn(rng, 64)),
                 DenseLayer(rng.normal(0, weight_sd * 5, (64, n_classes)).astype(np.float32), rng.normal(0, 0.02, n_classes).astype(np.float32), "none")]
    return net


# --------------------------------------------------------------------------- forward
def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even to bfloat16, returned as float64 (values exactly representable)."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).astype(np.float64)


def _act(x, act, alpha):
    if act == "relu":
        return np.maximum(x, 0.0)
    if act == "lrelu":
        return np.where(x >= 0.0, x, alpha * x)
    return x


def _bn_apply(x, bn, eps):
    if bn is None:
        return x
    g, b, m, v = (np.asarray(t, dtype=np.float64) for t in bn)
    return g * (x - m) / np.sqrt(v + eps) + b


def conv3x3_s2_same(x: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Keras Conv2D(3x3, strides 2, padding='same') on NHWC float64."""
    n, H, W, cin = x.shape
    Ho, Wo = (H + 1) // 2, (W + 1) // 2
    ph = max((Ho - 1) * 2 + 3 - H, 0)
    pw = max((Wo - 1) * 2 + 3 - W, 0)
    xp = np.pad(x, ((0, 0), (ph // 2, ph - ph // 2), (pw // 2, pw - pw // 2), (0, 0)))
    out = np.zeros((n, Ho, Wo, w.shape[3]), dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    for kh in range(3):
        for kw in range(3):
            patch = xp[:, kh:kh + 2 * Ho:2, kw:kw + 2 * Wo:2, :]
            out += patch @ w[kh, kw]
    return out + np.asarray(b, dtype=np.float64)


def _fold(layer, eps):
    w = np.asarray(layer.w, dtype=np.float64)
    b = np.asarray(layer.b, dtype=np.float64)
    if layer.bn is None:
        return w, b
    g, beta, m, v = (np.asarray(t, dtype=np.float64) for t in layer.bn)
    s = g / np.sqrt(v + eps)
    return w * s, (b - m) * s + beta


def tower_output(net: NetParams, X: np.ndarray, bf16_towers: bool = False) -> np.ndarray:
    """Conv towers only: (n, R, R, 3) -> float64 (n, 3, h, w, c) in the device's
    [branch][h][w][c] order.  bf16_towers=True reproduces the rounding points of the
    tensor-core towers: every layer's output activation is rounded to bf16 and the kernels of
    the layers after the first (BatchNorm folded, as on the device) are rounded to bf16."""
    X = np.asarray(X, dtype=np.float64)
    outs = []
    for br in range(3):
        h = X[:, :, :, br:br + 1]
        for li, layer in enumerate(net.branches[br]):
            if bf16_towers:
                w, b = _fold(layer, net.bn_eps)
                if li >= 1:
                    w = bf16_round(w)
                h = _act(conv3x3_s2_same(h, w, b), layer.act, net.alpha)
                if li + 1 < len(net.branches[br]):
                    h = bf16_round(h)
            else:
                h = conv3x3_s2_same(h, layer.w, layer.b)
                h = _act(_bn_apply(h, layer.bn, net.bn_eps), layer.act, net.alpha)
        outs.append(h)
    return np.stack(outs, axis=1)


def dense_from_tower(net: NetParams, tower: np.ndarray):
    """Dense stack + head evaluated in float64 on a given (already bf16-valued) tower output
    (n, 3, h, w, c), with the first kernel rounded to bf16 like the CUDA path."""
    n = tower.shape[0]
    fv = np.asarray(tower, dtype=np.float64).transpose(0, 2, 3, 1, 4).reshape(n, -1)   # Keras Flatten order
    d0 = net.dense[0]
    w1 = np.asarray(d0.w, dtype=np.float64)
    if d0.bn is not None:
        g, b, m, v = (np.asarray(t, dtype=np.float64) for t in d0.bn)
        s = g / np.sqrt(v + net.bn_eps)
        h = fv @ bf16_round(w1 * s) + ((np.asarray(d0.b, np.float64) - m) * s + b)
    else:
        h = fv @ bf16_round(w1) + np.asarray(d0.b, np.float64)
    h = _act(h, d0.act, net.alpha)
    for d in net.dense[1:]:
        h = _act(_bn_apply(h @ np.asarray(d.w, np.float64) + np.asarray(d.b, np.float64), d.bn, net.bn_eps),
                 d.act, net.alpha)
    if net.kind == "sgan_d":
        z = np.exp(h).sum(axis=1, keepdims=True)
        return z / (z + 1.0), h
    e = np.exp(h - h.max(axis=1, keepdims=True))
    return e / e.sum(axis=1, keepdims=True), h


def bf16_bits_to_float(bits: np.ndarray) -> np.ndarray:
    return (np.asarray(bits).astype(np.uint16).astype(np.uint32) << 16).view(np.float32).astype(np.float64)


def forward_bf16_towers(net: NetParams, X: np.ndarray):
    """Whole network with every rounding point of the tensor-core path (towers + dense)."""
    t = tower_output(net, X, bf16_towers=True)
    return dense_from_tower(net, bf16_round(t))


def forward(net: NetParams, X: np.ndarray, bf16_points: bool = False):
    """X (n, R, R, 3) float32 as ``preprocess`` returns -> (proba (n, C) float64, logits).

    bf16_points=True rounds exactly where the CUDA path does — the flattened conv-tower
    output and the first dense kernel — so the tcgen05 bf16 dense stack can be held to 1e-5;
    False is the pure float64 restatement of the Keras float32 graph."""
    X = np.asarray(X, dtype=np.float64)
    feats = []
    for br in range(3):
        h = X[:, :, :, br:br + 1]
        for layer in net.branches[br]:
            h = conv3x3_s2_same(h, layer.w, layer.b)
            h = _act(_bn_apply(h, layer.bn, net.bn_eps), layer.act, net.alpha)
        feats.append(h)
    fv = np.concatenate(feats, axis=3)              # concat on channels, xz|yz|xy
    fv = fv.reshape(fv.shape[0], -1)                # Flatten: (H, W, 96) row-major
    w1 = np.asarray(net.dense[0].w, dtype=np.float64)
    if bf16_points:
        fv = bf16_round(fv)
        d0 = net.dense[0]
        if d0.bn is not None:   # the CUDA path folds BN into the kernel before rounding it
            g, b, m, v = (np.asarray(t, dtype=np.float64) for t in d0.bn)
            s = g / np.sqrt(v + net.bn_eps)
            h = fv @ bf16_round(w1 * s) + ((np.asarray(d0.b, np.float64) - m) * s + b)
        else:
            h = fv @ bf16_round(w1) + np.asarray(d0.b, np.float64)
        h = _act(h, d0.act, net.alpha)
    else:
        d0 = net.dense[0]
        h = _act(_bn_apply(fv @ w1 + np.asarray(d0.b, np.float64), d0.bn, net.bn_eps), d0.act, net.alpha)
    for d in net.dense[1:]:
        h = _act(_bn_apply(h @ np.asarray(d.w, np.float64) + np.asarray(d.b, np.float64), d.bn, net.bn_eps),
                 d.act, net.alpha)
    logits = h
    if net.kind == "sgan_d":
        z = np.exp(logits).sum(axis=1, keepdims=True)       # sgan.py:125-129
        return z / (z + 1.0), logits
    e = np.exp(logits - logits.max(axis=1, keepdims=True))
    return e / e.sum(axis=1, keepdims=True), logits
