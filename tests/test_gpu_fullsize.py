"""-m gpu: BASELINE.json configs[1] at FULL size (65 536 cubes, 31.5 GB) through size-independent
properties, since the oracle cannot score that many scans in test time:
  * the fused K1||K2 pipeline and the serial K1 -> K2 order give bit-identical outputs
  * a scan's result does not depend on its batch: any gathered subset re-scored alone is identical
  * a 256-scan sample agrees with the CPU oracle (labels exact, probabilities 1e-5)
  * every probability row sums to 1 and labels are the row arg-max."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config1_full_size_properties(small_problem):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    free, _ = torch.cuda.mem_get_info()
    B = 65536
    if free < 40e9:
        pytest.skip("needs 40 GB of free HBM")
    sys.path.insert(0, ROOT)
    from bench import device_cubes
    from oracle import restate
    from radar_ml_b200.engine import Engine
    from radar_ml_b200.model import from_sklearn
    eng = Engine(0)
    eng.load_model(from_sklearn(small_problem["cal"]))
    cubes = device_cubes(B, 99, eng.device)
    # fused pipeline (default for large batches)
    eng.lib.rml_enable_timing(eng.ctx, 1)
    p1, l1, k1 = (t.clone() for t in eng.predict(cubes))
    eng.check_status()
    import ctypes as C
    fused = C.c_int()
    assert eng.lib.rml_last_timing(eng.ctx, None, None, C.byref(fused)) == 0 and fused.value == 1
    # serial order
    eng.lib.rml_set_fused(eng.ctx, 0, 0, 0)
    p2, l2, k2 = (t.clone() for t in eng.predict(cubes))
    eng.check_status()
    assert eng.lib.rml_last_timing(eng.ctx, None, None, C.byref(fused)) == 0 and fused.value == 0
    assert torch.equal(p1, p2) and torch.equal(l1, l2) and torch.equal(k1, k2)
    eng.lib.rml_set_fused(eng.ctx, 1, 0, 0)
    # batch-composition independence on a gathered, permuted subset (ragged size, below the fused threshold)
    g = torch.Generator(device="cpu").manual_seed(3)
    idx = torch.randperm(B, generator=g)[:4099].to(eng.device)
    sub = cubes[idx].contiguous()
    p3, l3, k3 = eng.predict(sub)
    eng.check_status()
    assert torch.equal(p3, p1[idx]) and torch.equal(l3, l1[idx]) and torch.equal(k3, k1[idx])
    # structural properties over the whole batch
    P = p1.double()
    assert float((P.sum(dim=1) - 1.0).abs().max()) < 1e-6
    assert torch.equal(P.argmax(dim=1).int(), l1)
    assert torch.equal((P.max(dim=1).values >= 0.7).to(torch.uint8), k1) or \
        float(((P.max(dim=1).values - 0.7).abs() < 1e-6).sum()) > 0      # fp32 copy of an fp64 compare
    # oracle on a sample
    sample = idx[:256]
    _, lab_o, _, known_o, P_o = restate.scan_path(cubes[sample].cpu().numpy(), small_problem["params"], mode="max")
    assert np.array_equal(l1[sample].cpu().numpy(), lab_o)
    assert np.abs(p1[sample].cpu().numpy().astype(np.float64) - P_o).max() < 1e-5
    eng.close()


@pytest.mark.parametrize("kind,n", [("dnn", 20000), ("sgan_c", 5000)])
def test_network_configs_size_independent_properties(kind, n):
    """configs[2] / configs[4] at sizes the float64 oracle cannot score in test time: the result of a
    scan must not depend on how the batch is cut into tower chunks and dense groups (the persistent
    tower kernel hands images between CTAs, roles and ring slots in a different order every time), nor
    on its position in the batch; rows are distributions; a sample agrees with the oracle."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    sys.path.insert(0, ROOT)
    from bench import device_cubes
    from oracle import nets, synth
    from radar_ml_b200.engine import Engine
    from radar_ml_b200.nets import GpuNetClassifier
    eng = Engine(0)
    spec = nets.random_dnn(5) if kind == "dnn" else nets.random_sgan(5)
    cubes = device_cubes(n, 41, eng.device)
    big = GpuNetClassifier(spec, engine=eng, chunk=9472 if kind == "dnn" else 4096)
    p1, l1 = (t.clone() for t in big.predict_cubes(cubes))
    small = GpuNetClassifier(spec, engine=eng, chunk=1000)             # ragged chunks, other CTA / image pairing
    p2, l2 = (t.clone() for t in small.predict_cubes(cubes))
    assert torch.equal(p1, p2) and torch.equal(l1, l2)
    g = torch.Generator(device="cpu").manual_seed(9)
    idx = torch.randperm(n, generator=g)[:777].to(eng.device)
    p3, l3 = small.predict_cubes(cubes[idx].contiguous())
    assert torch.equal(p3, p1[idx]) and torch.equal(l3, l1[idx])
    P = p1.double()
    assert float((P.sum(dim=1) - 1.0).abs().max()) < 1e-5 and bool(torch.isfinite(P).all())
    assert torch.equal(P.argmax(dim=1).int(), l1.int())
    # oracle on a few scans (float64 with the device's bf16 rounding points)
    m = 6
    sample = cubes[idx[:m]].cpu().numpy()
    xz, yz, xy = synth.project_max(sample)
    Xn = nets.preprocess([(xz[i], yz[i], xy[i]) for i in range(m)], spec.R)
    P_o, _ = nets.forward_bf16_towers(spec, Xn)
    assert np.abs(p1[idx[:m]].cpu().numpy() - P_o).max() < 5e-4
    eng.close()
