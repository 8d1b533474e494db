"""CPU tests of the host-side preparation in radar_ml_b200/nets.py (what is uploaded to the device):
bf16 rounding, BatchNorm folding, the Keras Flatten -> device order of the first Dense kernel."""
import numpy as np
import pytest


def test_to_bf16_bits_is_round_to_nearest_even():
    torch = pytest.importorskip("torch")
    from radar_ml_b200.nets import to_bf16_bits
    rng = np.random.default_rng(3)
    x = np.concatenate([
        rng.standard_normal(20000).astype(np.float32) * np.float32(10.0) ** rng.integers(-20, 20, 20000).astype(np.float32),
        np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, 1e-40, -1e-40, 3.3895314e38], dtype=np.float32),
        # exact ties between two bf16 neighbours (low 16 bits = 0x8000): even mantissa wins
        np.array([0x3F808000, 0x3F818000, 0xBF808000, 0x00008000, 0x7F7F8000], dtype=np.uint32).view(np.float32),
    ])
    want = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(to_bf16_bits(x), want)


def test_fold_bn_equals_conv_then_batchnorm():
    """y = BN(conv(x)) with inference statistics == conv'(x) with folded kernel and bias (sgan.py:136-141),
    and identical to the oracle's own folding."""
    from oracle import nets as onets
    from radar_ml_b200.nets import fold_bn
    rng = np.random.default_rng(5)
    w = rng.standard_normal((3, 3, 4, 6))
    b = rng.standard_normal(6)
    bn = (rng.uniform(0.5, 1.5, 6), rng.standard_normal(6), rng.standard_normal(6), rng.uniform(0.1, 2.0, 6))
    eps = 1e-3
    wf, bf = fold_bn(w, b, bn, eps)
    x = rng.standard_normal((5, 4))                       # one 1x1 "pixel" per row, tap (1,1) only
    y = x @ w[1, 1] + b
    g, beta, m, v = bn
    want = g * (y - m) / np.sqrt(v + eps) + beta
    assert np.allclose(x @ wf[1, 1] + bf, want, rtol=0, atol=1e-12)
    layer = onets.ConvLayer(w=w, b=b, act="lrelu", bn=bn)
    wo, bo = onets._fold(layer, eps)
    assert np.allclose(wf, wo, rtol=0, atol=1e-15) and np.allclose(bf, bo, rtol=0, atol=1e-15)
    w0, b0 = fold_bn(w, b, None, eps)
    assert np.array_equal(w0, w) and np.array_equal(b0, b)


def test_dense1_permutation_keras_flatten_to_device_order():
    """Keras Flatten of concat([xz, yz, xy], axis=-1) indexes (h, w, branch*cl + c) (dnn.py:76-79); the
    device stores the tower output as [branch][h][w][c].  The permutation used at load time must send
    row k of the Keras kernel to the device row holding the same activation."""
    hw, cl = 4, 3
    K = 3 * hw * hw * cl
    w1 = np.arange(K * 2, dtype=np.float64).reshape(K, 2)               # row index recoverable from column 0
    dev = w1.reshape(hw, hw, 3, cl, 2).transpose(2, 0, 1, 3, 4).reshape(K, 2)   # as in GpuNetClassifier._load
    for br in range(3):
        for h in range(hw):
            for w in range(hw):
                for c in range(cl):
                    keras_row = (h * hw + w) * (3 * cl) + br * cl + c
                    dev_row = ((br * hw + h) * hw + w) * cl + c
                    assert dev[dev_row, 0] == w1[keras_row, 0]
