"""CPU tests of the network oracle (oracle/nets.py): Pillow parity of the restated bicubic
resize, architecture known-answers from the reference's model diagrams / survey, and an
independent torch check of the restated Keras layers."""
import numpy as np
import pytest

from oracle import nets


@pytest.mark.parametrize("shape", [(22, 176), (31, 176), (22, 31)])
@pytest.mark.parametrize("R", [80, 128])
def test_resize_restatement_is_bit_exact_to_pillow(shape, R):
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(shape[0] * R)
    p = ((rng.integers(0, 256, shape).astype(np.float32)) - 127.5) / 127.5
    p = p.astype(np.float32)
    ref = np.asarray(Image.fromarray(p).resize((R, R), resample=Image.BICUBIC))   # dnn.py:243
    got = nets.pil_bicubic_resize(p, R)
    assert ref.dtype == got.dtype == np.float32 and np.array_equal(ref, got)


def test_host_tables_match_oracle_tables():
    from radar_ml_b200.nets import pil_bicubic_tables
    for n_in, n_out in ((176, 80), (22, 80), (31, 128), (176, 128)):
        K1, B1 = nets.pil_coeffs(n_in, n_out)
        K2, B2 = pil_bicubic_tables(n_in, n_out)
        assert np.array_equal(K1, K2) and np.array_equal(B1, B2)
    K, B = nets.pil_coeffs(176, 80)
    assert K.shape[1] == 11 and B[:, 1].max() <= 11        # <= 9-11 taps when shrinking 2.2x
    assert nets.pil_coeffs(22, 80)[0].shape[1] == 5         # 4 taps (+1) when enlarging


def test_parameter_counts_match_reference_models():
    d = nets.random_dnn(0)
    n = sum(l.w.size + l.b.size for br in d.branches for l in br) + sum(l.w.size + l.b.size for l in d.dense)
    assert n == 2519331                       # SURVEY.md §8a A13 (images/dnn_model.png)
    s = nets.random_sgan(0)
    n = sum(l.w.size + 5 * l.b.size for br in s.branches for l in br) + \
        sum(l.w.size + l.b.size + (4 * l.b.size if l.bn else 0) for l in s.dense)
    assert n == 1861091                       # SURVEY.md §8a A14 (images/sgan_c_model.png)
    assert d.dense[0].w.shape[0] == 38400 and s.dense[0].w.shape[0] == 24576


def test_conv_same_stride2_matches_torch():
    torch = pytest.importorskip("torch")
    F = torch.nn.functional
    rng = np.random.default_rng(1)
    for (H, cin, cout) in ((80, 1, 64), (40, 64, 32), (16, 8, 32)):
        x = rng.normal(size=(2, H, H, cin))
        w = rng.normal(size=(3, 3, cin, cout)).astype(np.float32)
        b = rng.normal(size=(cout,)).astype(np.float32)
        got = nets.conv3x3_s2_same(x, w, b)
        xt = F.pad(torch.from_numpy(x).permute(0, 3, 1, 2), (0, 1, 0, 1))      # TF 'same': pad after
        ref = F.conv2d(xt, torch.from_numpy(w.astype(np.float64)).permute(3, 2, 0, 1),
                       torch.from_numpy(b.astype(np.float64)), stride=2).permute(0, 2, 3, 1).numpy()
        assert got.shape == (2, H // 2, H // 2, cout) and np.abs(got - ref).max() < 1e-12


def test_forward_heads_and_bf16_points():
    rng = np.random.default_rng(2)
    X = rng.uniform(-1, 1, size=(3, 80, 80, 3)).astype(np.float32)
    d = nets.random_dnn(1)
    P, lg = nets.forward(d, X)
    assert P.shape == (3, 3) and np.allclose(P.sum(axis=1), 1.0)
    Pb, _ = nets.forward(d, X, bf16_points=True)
    assert 0 < np.abs(P - Pb).max() < 5e-3     # bf16 rounding is visible but small
    X = rng.uniform(-1, 1, size=(2, 128, 128, 3)).astype(np.float32)
    c = nets.random_sgan(1, kind="sgan_c")
    dmod = nets.random_sgan(1, kind="sgan_d")
    Pc, lc = nets.forward(c, X)
    Pd, ld = nets.forward(dmod, X)
    assert np.allclose(lc, ld)                 # shared trunk (sgan.py:205 vs 210)
    z = np.exp(ld).sum(axis=1, keepdims=True)
    assert Pd.shape == (2, 1) and np.allclose(Pd, z / (z + 1))


def test_preprocess_layout():
    rng = np.random.default_rng(3)
    samples = [(rng.integers(0, 256, (22, 176)).astype(np.float32),
                rng.integers(0, 256, (31, 176)).astype(np.float32),
                rng.integers(0, 256, (22, 31)).astype(np.float32)) for _ in range(2)]
    X = nets.preprocess(samples, 80)
    assert X.shape == (2, 80, 80, 3) and X.dtype == np.float32
    Image = pytest.importorskip("PIL.Image")
    q = ((samples[1][2] - 127.5) / 127.5).astype(np.float32)
    assert np.array_equal(X[1, :, :, 2], np.asarray(Image.fromarray(q).resize((80, 80), resample=Image.BICUBIC)))


def _torch_graph(net, X):
    """dnn.py:55-91 / sgan.py:157-217 written a second time, independently of oracle/nets.py, with
    torch.nn modules in float64: Conv2d on TF-'same'-padded NCHW, BatchNorm in eval mode with the
    moving statistics (eps 1e-3), LeakyReLU(0.2) / ReLU, channel concat, Keras Flatten (H, W, C),
    Linear + BatchNorm1d.  Returns (probabilities, logits, per-layer output shapes)."""
    import torch
    nn = torch.nn
    t = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64))  # noqa: E731
    act = {"relu": nn.ReLU(), "lrelu": nn.LeakyReLU(net.alpha), "none": nn.Identity()}
    shapes, feats = [], []
    with torch.no_grad():
        for br in range(3):
            h = t(X[:, :, :, br:br + 1]).permute(0, 3, 1, 2)            # NHWC -> NCHW
            for layer in net.branches[br]:
                cin, cout = layer.w.shape[2], layer.w.shape[3]
                conv = nn.Conv2d(cin, cout, 3, stride=2).double()
                conv.weight.copy_(t(layer.w).permute(3, 2, 0, 1))        # HWIO -> OIHW
                conv.bias.copy_(t(layer.b))
                h = conv(nn.functional.pad(h, (0, 1, 0, 1)))              # TF 'same', stride 2: pad after
                if layer.bn is not None:
                    g, b, m, v = layer.bn
                    bn = nn.BatchNorm2d(cout, eps=net.bn_eps).double().eval()
                    bn.weight.copy_(t(g)); bn.bias.copy_(t(b)); bn.running_mean.copy_(t(m)); bn.running_var.copy_(t(v))
                    h = bn(h)
                h = act[layer.act](h)
                shapes.append(tuple(h.shape[1:]))
            feats.append(h)
        h = torch.cat(feats, dim=1)                                       # concat on channels xz | yz | xy
        shapes.append(tuple(h.shape[1:]))
        h = h.permute(0, 2, 3, 1).reshape(h.shape[0], -1)                 # Keras Flatten: (H, W, C) row-major
        shapes.append(tuple(h.shape[1:]))
        for d in net.dense:
            lin = nn.Linear(d.w.shape[0], d.w.shape[1]).double()
            lin.weight.copy_(t(d.w).T); lin.bias.copy_(t(d.b))
            h = lin(h)
            if d.bn is not None:
                g, b, m, v = d.bn
                bn = nn.BatchNorm1d(d.w.shape[1], eps=net.bn_eps).double().eval()
                bn.weight.copy_(t(g)); bn.bias.copy_(t(b)); bn.running_mean.copy_(t(m)); bn.running_var.copy_(t(v))
                h = bn(h)
            h = act[d.act](h)
        logits = h
        if net.kind == "sgan_d":
            z = torch.exp(logits).sum(dim=1, keepdim=True)
            return (z / (z + 1)).numpy(), logits.numpy(), shapes
        return torch.softmax(logits, dim=1).numpy(), logits.numpy(), shapes


@pytest.mark.parametrize("kind", ["dnn", "sgan_c", "sgan_d"])
def test_whole_graph_against_independent_torch_modules(kind):
    """Keras is not installable here (parity unpinned against TensorFlow itself), so the restated
    graph is at least held to a second, independently written implementation on torch.nn layers,
    and to the layer shapes printed in the reference's model diagrams."""
    pytest.importorskip("torch")
    rng = np.random.default_rng(11)
    if kind == "dnn":
        net = nets.random_dnn(5)
        X = rng.uniform(-1, 1, size=(3, 80, 80, 3)).astype(np.float32)
        # images/dnn_model.png: 80 -> 40x40x64 -> 20x20x32 per branch, concat 20x20x96, flatten 38400
        want = [(64, 40, 40), (32, 20, 20)] * 3 + [(96, 20, 20), (38400,)]
    else:
        net = nets.random_sgan(5, kind=kind)
        X = rng.uniform(-1, 1, size=(2, 128, 128, 3)).astype(np.float32)
        # images/sgan_c_model.png: 128 -> 64x64x128 -> 32x32x64 -> 16x16x32, concat 16x16x96, flatten 24576
        want = [(128, 64, 64), (64, 32, 32), (32, 16, 16)] * 3 + [(96, 16, 16), (24576,)]
    P_t, lg_t, shapes = _torch_graph(net, X)
    assert shapes == want
    P_o, lg_o = nets.forward(net, X)
    assert np.abs(lg_o - lg_t).max() < 1e-9 and np.abs(P_o - P_t).max() < 1e-10


def test_trained_magnitude_weights_keep_the_rounding_points_tight():
    """Random-init weights are small; a trained network has O(1) activations and saturated
    softmax outputs.  The bf16 rounding points of the device path must stay close to the float64
    graph for such weights too (this is what bounds the GPU tests' tolerances)."""
    rng = np.random.default_rng(21)
    net = nets.random_dnn(2)
    for br in net.branches:
        for layer in br:
            layer.w = (layer.w * 3.0).astype(np.float32)          # larger kernels -> O(1..10) activations
    net.dense[2].w = (net.dense[2].w * 20.0).astype(np.float32)   # confident logits
    X = rng.uniform(-1, 1, size=(4, 80, 80, 3)).astype(np.float32)
    P64, lg64 = nets.forward(net, X)
    Pbf, lgbf = nets.forward_bf16_towers(net, X)
    assert np.abs(lg64).max() > 1.0                                # the logits are no longer tiny
    assert np.abs(Pbf - P64).max() < 2e-2 and np.array_equal(Pbf.argmax(1), P64.argmax(1))
