#!/usr/bin/env python
"""bench.py — radar scans/sec (projection + classify) on N B200s, with roofline + CPU baseline.

A "step" is one pass of the hot path (K1 projection -> K2 SVC-RBF scoring [-> label
all-gather when N > 1]) over one batch of synthetic cubes already resident in HBM.
  N = 1 : BASELINE.json configs[1]  — 65 536 cubes (31.47 GB), SVC-RBF 3-class
  N > 1 : BASELINE.json configs[3]  — 131 072 cubes per GPU (1 M at N = 8), weak scaling
Launch: ``python bench.py --gpus 1`` or, for N > 1,
``python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1
--master-port P bench.py --gpus N``.
``--impl reference`` times the reference's own per-scan CPU path (oracle port driving
scikit-learn's libsvm, all host cores) on the same workload definition.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CUBE_BYTES = 22 * 31 * 176 * 4            # 480 128 B, common.py:25-27
ALGO_BYTES_PER_SCAN = CUBE_BYTES + 12 + 4  # + 3 fp32 probs + int32 label
# dram__bytes_read.sum + dram__bytes_write.sum of k1_project_max<u8> per scan, from the round-2 ncu
# --set full capture summarised in profiles/r2_k1_project_max_u8_ncu.txt (7.867382 GB read +
# 0.122009 GB written for 16 384 scans; round 1: 7.86665 + 0.12745): 1.016 x the algorithmic bytes,
# i.e. no re-reads.
K1_DRAM_TRAFFIC_PER_SCAN = (7.867382e9 + 122.009088e6) / 16384
METRIC = "radar_scans_per_sec_proj_classify"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries exactly ONE line, the JSON result.  Native libraries write to file descriptor 1
# behind Python's back (NCCL prints "NCCL version ..." there even with NCCL_DEBUG_FILE set), so fd 1
# is pointed at stderr for the whole run and the result goes out through a private copy of the
# original stdout.
_RESULT_OUT = None


def claim_stdout():
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, torch copy read+write)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def device_cubes(n, seed, device, chunk=1024, integer=True):
    """Integer-valued sparse-blob cubes generated ON the device: one anisotropic
    Gaussian blob per scan, amplitude U(.5,1)*255, N(0,6) noise, rint, <13 -> 0, clip [0,255]."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((n, 22, 31, 176), device=device, dtype=torch.float32)
    gi = torch.arange(22, device=device, dtype=torch.float32).view(1, 22, 1, 1)
    gj = torch.arange(31, device=device, dtype=torch.float32).view(1, 1, 31, 1)
    gk = torch.arange(176, device=device, dtype=torch.float32).view(1, 1, 1, 176)
    sig = torch.tensor([[1.6, 2.2, 5.0], [2.4, 3.4, 8.0], [3.6, 6.0, 13.0]], device=device)
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        cls = torch.randint(0, 3, (m,), device=device, generator=g)
        u = torch.rand((m, 8), device=device, generator=g)
        s = sig[cls] * (0.8 + 0.45 * u[:, 0:3])
        ci = (2 + u[:, 3] * 17).view(m, 1, 1, 1)
        cj = (3 + u[:, 4] * 24).view(m, 1, 1, 1)
        ck = (10 + u[:, 5] * 155).view(m, 1, 1, 1)
        amp = ((0.5 + 0.5 * u[:, 6]) * 255.0).view(m, 1, 1, 1)
        v = amp * torch.exp(-0.5 * (((gi - ci) / s[:, 0].view(m, 1, 1, 1)) ** 2
                                    + ((gj - cj) / s[:, 1].view(m, 1, 1, 1)) ** 2
                                    + ((gk - ck) / s[:, 2].view(m, 1, 1, 1)) ** 2))
        v += 6.0 * torch.randn(v.shape, device=device, generator=g)
        if integer:
            v = torch.round(v)
            v[v < 13.0] = 0.0
        out[lo:lo + m] = v.clamp_(0.0, 255.0)      # integer=False: real-valued scans, still in range
        del v
    return out


class ClockSampler:
    """SM clock + throttle reasons sampled through NVML every ~5 ms DURING the timed region
    (the recipe's nvidia-smi line needs >= 100 ms per sample; a timed region here is ~50 ms)."""
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
            0x4: "sw_power_cap"}

    def __init__(self, index):
        import threading
        self.sm, self.reasons, self.power = [], set(), []
        self.max = None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
            self.max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        except Exception as e:  # pragma: no cover
            log("[bench] NVML unavailable: %r" % (e,))

    def _run(self):
        nv = self.nv
        it = 0
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                if it % 8 == 0:      # the reasons / power queries are slower than the clock query
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                    for bit, name in self.BITS.items():
                        if r & bit:
                            self.reasons.add(name)
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            it += 1
            time.sleep(0.004)      # every NVML query takes the driver's lock: poll gently (a 1 ms poll cost the
                                   # timed region up to 4 % on a box whose NVML calls took 2.5 ms each)

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=2)
        out = {"sm_mhz": None, "sm_max_mhz": self.max, "reasons": sorted(self.reasons),
               "samples": len(self.sm),
               "window": ("timed region + the same steps repeated untimed until 16 samples"
                          if getattr(self, "extended", False) else "timed region")}
        if self.sm:
            out["sm_mhz"] = float(np.median(self.sm))
        if self.power:
            out["power_w_max"] = float(max(self.power))
        return out


def build_model(seed=1234):
    """SVC-RBF C=10 gamma=0.01 (train_svc.log:24-31) on the reference's split sizes
    (909 train / 114 val, train_svc.log:11-13), synthetic MAX-projection features."""
    from oracle import synth
    n_train = int(os.environ.get("RML_BENCH_TRAIN", "909"))     # test hook: a smaller fit
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        n_val = 114 if n_train == 909 else max(30, n_train // 8)      # train_svc.log:11-13 split sizes
        return synth.standard_model(n_train=n_train, n_val=n_val, seed=seed, mode="max")


# ------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline, synth
    cores = cpu_baseline.host_cores()
    per_gpu = args.scans_per_gpu or (65536 if args.gpus == 1 else 131072)
    sample = int(os.environ.get("RML_BENCH_SAMPLE", "4096"))   # bounded CPU sample per step
    log("[reference] fitting model, generating %d sample scans on the host" % sample)
    cal = build_model()
    cubes, _, _ = synth.make_cubes(sample, seed=4321)
    classes = synth.CLASSES
    for _ in range(args.warmup):
        cpu_baseline.run_all_cores(cubes[: max(cores, 8)], cal, classes, cores=cores)
    v1, _, lat1 = cpu_baseline.run_one_core(cubes[:48], cal, classes, repeats=1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_baseline.run_all_cores(cubes, cal, classes, cores=cores)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "scans/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "predict.py per-scan loop (MAX projections + process_samples + "
                               "CalibratedClassifierCV(SVC-RBF).predict_proba), all host cores",
                   "scans_per_gpu": per_gpu, "sample_scans_per_step": sample,
                   "n_sv": int(cal.calibrated_classifiers_[0].estimator.estimator.support_vectors_.shape[0])},
        "cpu_baseline": {"value": value, "unit": "scans/s", "cores": cores, "kind": "port",
                         "cpu_model": cpu_baseline.cpu_model_string(),
                         "sample": "%d scans/step, one worker process per core, sklearn libsvm, ndimage.zoom "
                                   "called like common.py:143" % sample,
                         "one_core": {"value": v1, "unit": "scans/s", "ms_per_scan_median": lat1}},
        "e2e": {"value": value, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------ our arm
DNN_FLOP = 54690176        # SURVEY.md §8a A13, per scan
SGAN_FLOP = 512762240      # A14


def tensor_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1400.0, "fallback (B200_PROFILING.md ~1.4 PFLOP/s sustained)"


def bind_to_gpu_node(local, world):
    """Pin this rank's threads (and therefore its first-touch / pinned host memory) to the CPUs
    NVML reports as local to its GPU.  When every GPU reports the same set (a single-node guest)
    the set is split evenly so that the ranks' copy threads do not share cores."""
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[local]) if vis and vis.split(",")[local].isdigit() else local
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = sorted(cpus & allowed) or sorted(allowed)
        info["gpu_local_cpus"] = "%d-%d (%d)" % (cpus[0], cpus[-1], len(cpus))
        try:
            info["numa_node"] = int(pynvml.nvmlDeviceGetNumaNodeId(h))
        except Exception:
            pass
        if world > 1 and len(cpus) >= 2 * world:
            per = len(cpus) // world
            cpus = cpus[local * per:(local + 1) * per]
        os.sched_setaffinity(0, cpus)
        info["bound"] = True
        info["cpus"] = "%d-%d" % (cpus[0], cpus[-1])
    except Exception as e:  # pragma: no cover
        info["error"] = repr(e)
    return info


def label_checksum(t):
    """Order-sensitive 62-bit checksum of an int32 label vector, computed on the device."""
    import torch
    idx = torch.arange(1, t.numel() + 1, device=t.device, dtype=torch.int64)
    return int(((t.to(torch.int64) + 1) * (idx % 1000003 + 1)).sum().item() % (1 << 62))


def net_leg(kind, eng_dev, cubes, n_scans, passes, steps, rank, world, dist, stream):
    """configs[2] (dnn.py forward) / configs[4] (sgan.py c_model forward) from resident cubes."""
    import torch
    from oracle import nets as onets
    from radar_ml_b200.engine import Engine
    from radar_ml_b200.nets import GpuNetClassifier
    dev = cubes.device
    eng = Engine(dev.index)
    spec = onets.random_dnn(0) if kind == "dnn" else onets.random_sgan(0)
    net = GpuNetClassifier(spec, engine=eng, chunk=int(os.environ.get("RML_BENCH_NET_CHUNK", "9472" if kind == "dnn" else "4096")))   # dnn: half a dense group (148 x 128 scans): equal launches
    sub = cubes[:n_scans]
    launches0 = eng.launch_count
    out = net.predict_cubes(sub)
    torch.cuda.synchronize(dev)
    launches_per_pass = eng.launch_count - launches0
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        for _ in range(passes):
            out = net.predict_cubes(sub)
    e1.record(stream)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.cpu()[0])
    total = world * n_scans * passes
    value = total / ms * 1e3
    peak_h, _ = measured_peaks()
    peak_t, peak_t_src = tensor_peak()
    flop = DNN_FLOP if kind == "dnn" else SGAN_FLOP
    leg = {"workload": ("configs[2]: dnn.py forward (3 towers Conv64-Conv32 -> Dense 64-64-3) on %d cubes "
                        "(%d resident cubes x %d passes), bf16 tensor cores, 1xB200" % (n_scans * passes, n_scans, passes)
                        if kind == "dnn" else
                        "configs[4]: sgan.py c_model forward (3 towers Conv128-64-32+BN+LeakyReLU -> Dense) on "
                        "%d cubes/GPU x %d GPUs, from raw cubes (K1 -> resize 128x128 -> towers -> dense)" % (n_scans, world)),
           "value": value, "unit": "scans/s", "ms_per_step": ms, "scans_per_step": total,
           "roofline": {"hbm_frac": (value / world) * CUBE_BYTES / (peak_h * 1e9),
                        "tensor_frac": (value / world) * flop / (peak_t * 1e12),
                        "tflops": (value / world) * flop / 1e12, "flop_per_scan": flop,
                        "tensor_peak_tflops": peak_t, "tensor_peak_source": peak_t_src},
           "gpu_launches_per_pass": int(launches_per_pass), "igemm": bool(net.uses_igemm)}
    if rank == 0:
        # parity of a sample against the CPU restatement (same bf16 rounding points), not timed
        n_par = 12
        from oracle import synth
        c_np = sub[:n_par].cpu().numpy()
        xz, yz, xy = synth.project_max(c_np)
        X = onets.preprocess([(xz[i], yz[i], xy[i]) for i in range(n_par)], spec.R)
        P_o, _ = onets.forward_bf16_towers(spec, X)
        P_g = out[0][:n_par].cpu().numpy().astype(np.float64)
        srt = np.sort(P_o, axis=1)
        clear = (srt[:, -1] - srt[:, -2]) > 1e-3
        leg["parity"] = {"scans": n_par, "max_abs_dproba": float(np.abs(P_g - P_o).max()),
                         "labels_equal": bool(np.array_equal(out[1][:n_par].cpu().numpy()[clear],
                                                             np.argmax(P_o, axis=1)[clear])),
                         "oracle": "oracle/nets.py float64 with the device's bf16 rounding points (parity unpinned: "
                                   "no Keras in the image)"}
    eng.close()
    return leg


def run_ours(args):
    import torch
    import torch.distributed as dist
    from oracle import cpu_baseline, restate, synth
    from radar_ml_b200.engine import Engine
    from radar_ml_b200.model import from_sklearn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)" % (args.gpus, world))
    binding = bind_to_gpu_node(local, world)      # before any pinned allocation
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    B = args.scans_per_gpu or (65536 if world == 1 else 131072)
    # N = 1 also holds config 4's per-GPU shard (131 072 scans) so that the 1 -> 8 GPU curve can be
    # read like for like (`shard_131072`); the headline at N = 1 stays configs[1] = 65 536 scans
    B_alloc = B if (args.scans_per_gpu or world > 1) else 131072
    cal = build_model()
    params = from_sklearn(cal)
    eng = Engine(local)
    eng.load_model(params)
    assert eng.model_is_integral, "bench model must take the u8 tensor-core path"
    if rank == 0:
        log("[bench] model n_sv=%d; generating %d cubes (%.2f GB) per GPU" % (params.n_sv, B_alloc, B_alloc * CUBE_BYTES / 1e9))
    cubes_all = device_cubes(B_alloc, 1234 + rank, dev)
    cubes = cubes_all[:B]
    C = params.n_classes
    proba = torch.empty((B_alloc, C), device=dev, dtype=torch.float32)
    known = torch.empty((B_alloc,), device=dev, dtype=torch.uint8)
    # the scorer writes its labels straight into this rank's slice of the gather buffer
    gathered = torch.empty((world * B_alloc,), device=dev, dtype=torch.int32)
    label = gathered[rank * B:(rank + 1) * B] if world > 1 else gathered[:B]
    stream = torch.cuda.current_stream(dev)

    import ctypes as Ct
    lib, ctx = eng.lib, eng.ctx
    sp = Ct.c_void_p(stream.cuda_stream)
    if world > 1:
        # the exchange goes through the C ABI (rml_allgather_labels -> ncclAllGather); torch.distributed
        # only carries the 128-byte communicator id and the timing reductions
        box = [eng.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        eng.comm_init(rank, world, box[0])

    work = eng.workspace(B_alloc)
    lib.rml_enable_timing(ctx, 1)

    def local_step(n=B, lab=None):
        # the public device entry point: K1 and K2 as one pipeline (fused for large batches)
        lab = label if lab is None else lab
        rc = lib.rml_predict(ctx, Ct.c_void_p(cubes_all.data_ptr()), n, 0, None, 7, 0.7,
                             Ct.c_void_p(work.data_ptr()), Ct.c_void_p(proba.data_ptr()),
                             Ct.c_void_p(lab.data_ptr()), Ct.c_void_p(known.data_ptr()), sp)
        assert rc == 0, lib.rml_last_error(ctx)

    def step():
        local_step()
        if world > 1:
            rc = lib.rml_allgather_labels(ctx, Ct.c_void_p(label.data_ptr()), Ct.c_void_p(gathered.data_ptr()), B, sp)
            assert rc == 0, lib.rml_last_error(ctx)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    eng.check_status()
    barrier()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = eng.launch_count
    barrier()
    e0.record(stream)
    for s in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    launches = eng.launch_count - launches0 + (args.steps if world > 1 else 0)
    ms = e0.elapsed_time(e1)
    if sampler is not None and len(sampler.sm) < 16:
        # NVML queries take tens of ms on some boxes, so a ~100 ms timed region yields only a few
        # samples: keep the same steps running (untimed) until the sampler has a usable median
        t_end = time.perf_counter() + 3.0
        while len(sampler.sm) < 16 and time.perf_counter() < t_end:
            local_step()      # rank 0 only: no collective in here
            torch.cuda.synchronize(dev)
        sampler.extended = True
    clocks = sampler.stop() if sampler else None
    # per-kernel durations: a second pass of the same K steps with the library's own CUDA events
    # around the projection kernel (reading them synchronises, so it is kept out of the timed loop)
    k1_list, tot_list = [], []
    k1v, totv, fusedv = Ct.c_float(), Ct.c_float(), Ct.c_int()
    for s in range(args.steps):
        step()
        rc = lib.rml_last_timing(ctx, Ct.byref(k1v), Ct.byref(totv), Ct.byref(fusedv))
        assert rc == 0, lib.rml_last_error(ctx)
        k1_list.append(k1v.value)
        tot_list.append(totv.value)
    k1_ms = float(np.mean(k1_list))
    k2_ms = float(np.mean(tot_list)) - k1_ms      # scorer time NOT hidden behind the projection
    fused = int(fusedv.value)
    if world > 1:
        t = torch.tensor([ms, k1_ms, k2_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, k1_ms, k2_ms = (float(x) for x in t.cpu())
    value = world * B * args.steps / (ms * 1e-3)

    if args.skip_extras:
        if rank == 0:
            emit({"metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": world,
                  "ms_per_step": ms / args.steps, "k1_ms": k1_ms, "k2_exposed_ms": k2_ms, "fused": bool(fused),
                  "note": "--skip-extras profiling run, not a bench line"})
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- multi-GPU determinism: the gathered vector is the same on every rank,
    # slice r is what rank r scored, and a non-zero rank's shard agrees with the CPU oracle
    step()
    eng.check_status()
    barrier()
    sums = [label_checksum(gathered[r * B:(r + 1) * B]) for r in range(world)] if world > 1 else [label_checksum(label)]
    gathered_equal, ranks_checked = True, 1
    if world > 1:
        mine = torch.tensor(sums, device=dev, dtype=torch.int64)
        allv = torch.empty((world, world), device=dev, dtype=torch.int64)
        dist.all_gather_into_tensor(allv, mine)
        allv = allv.cpu().numpy()
        gathered_equal = bool((allv == allv[0:1]).all())
        ranks_checked = world
    gather_par = {"gathered_equal": gathered_equal, "ranks_checked": ranks_checked,
                  "slice_checksums": [str(v) for v in sums],
                  "note": "slice r = the labels rank r scored for its seeded shard (seed 1234 + r): equal "
                          "checksums across the N = 1/2/4/8 lines mean identical labels element for element"}

    # ---- e2e: host buffers through the public C-ABI call, H2D + D2H inside the timed region
    Be = min(args.e2e_scans, B)
    host = torch.empty((Be, 22, 31, 176), dtype=torch.float32, pin_memory=True)   # allocated after the CPU bind
    host.copy_(cubes[:Be])
    torch.cuda.synchronize(dev)
    host_np = host.numpy()
    out = (np.empty((Be, C), np.float32), np.empty((Be,), np.int32), np.empty((Be,), np.uint8))
    # the library's default: float32 cubes holding the sensor's integers are narrowed to bytes by host
    # threads inside rml_predict_host (checked exact; it switches itself off where the host converts
    # slower than the bus moves float32) -- 3 warm-up calls so that this decision is made before timing
    for _ in range(3):
        eng.predict_host(host_np, mode="max", out=out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.predict_host(host_np, mode="max", out=out)
    e2e_local = time.perf_counter() - t0
    barrier()
    e2e_s = time.perf_counter() - t0
    xfer = eng.last_host_transfer()          # bytes that actually crossed the bus in the last call
    # the same call with the narrowing OFF: every float32 byte crosses the bus
    eng.set_host_narrowing(False)
    out_f = (np.empty((Be, C), np.float32), np.empty((Be,), np.int32), np.empty((Be,), np.uint8))
    for _ in range(2):
        eng.predict_host(host_np, mode="max", out=out_f)
    barrier()
    t0f = time.perf_counter()
    for _ in range(args.steps):
        eng.predict_host(host_np, mode="max", out=out_f)
    barrier()
    e2e_f32_s = time.perf_counter() - t0f
    xfer_f = eng.last_host_transfer()
    eng.set_host_narrowing(True)
    narrow_equal = bool(np.array_equal(out[1], out_f[1]) and np.array_equal(out[0], out_f[0]) and np.array_equal(out[2], out_f[2]))
    # the ceiling the host side sets: the same pinned buffer copied H2D by every rank at once, no kernels
    devbuf = torch.empty((min(Be, 2048), 22, 31, 176), device=dev, dtype=torch.float32)
    nb = devbuf.shape[0]
    devbuf.copy_(host[:nb], non_blocking=True)
    barrier()
    t1 = time.perf_counter()
    reps = max(2, (Be * args.steps) // nb // 2)
    for _ in range(reps):
        devbuf.copy_(host[:nb], non_blocking=True)
    torch.cuda.synchronize(dev)
    copy_local = time.perf_counter() - t1
    barrier()
    del devbuf
    h2d_rank = xfer["h2d_bytes"] * args.steps / e2e_local / 1e9          # bytes on the bus
    in_rank = Be * args.steps * CUBE_BYTES / e2e_local / 1e9               # float32 bytes consumed from host memory
    ceil_rank = reps * nb * CUBE_BYTES / copy_local / 1e9
    if world > 1:
        t = torch.tensor([e2e_s, -h2d_rank, -ceil_rank, h2d_rank, ceil_rank, e2e_f32_s], device=dev, dtype=torch.float64)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        e2e_s = float(tmax[0])
        e2e_f32_s = float(tmax[5])
        h2d_min, ceil_min = -float(tmax[1]), -float(tmax[2])
        h2d_sum, ceil_sum = float(tsum[3]), float(tsum[4])
    else:
        h2d_min = h2d_sum = h2d_rank
        ceil_min = ceil_sum = ceil_rank
    e2e_value = world * Be * args.steps / e2e_s
    e2e_labels = out[1].copy()

    # ---- parity of a seeded subset against the oracle (not timed)
    parity = None
    cpu = None
    latency = None
    if rank == 0:
        n_par = 192
        sub = cubes[:n_par].cpu().numpy()
        p = restate.export_params(cal)
        _, lab_o, _, known_o, P_o = restate.scan_path(sub, p, mode="max")
        lab_g = label[:n_par].cpu().numpy()
        parity = {"scans": n_par,
                  "labels_equal": bool(np.array_equal(lab_g, lab_o)),
                  "known_equal": bool(np.array_equal(known[:n_par].cpu().numpy().astype(bool), known_o)),
                  "max_abs_dproba": float(np.abs(proba[:n_par].cpu().numpy().astype(np.float64) - P_o).max()),
                  "e2e_labels_equal": bool(np.array_equal(e2e_labels[:n_par], lab_o))}
        parity.update(gather_par)
        sample = int(min(Be, 4096))
        # second, independent checker: the plain-C oracle (OpenMP over scans) on 4 096 scans of THIS
        # rank and, at N > 1, on 4 096 scans of the LAST rank's shard as they arrived through the gather
        try:
            from oracle import c_oracle
            cm = c_oracle.CModel(p)
            c_oracle.scan_path(host_np[:64], cm)
            t0 = time.perf_counter()
            _, lab_c, _, known_c, P_c = c_oracle.scan_path(host_np[:sample], cm, mode="max")
            dtc = time.perf_counter() - t0
            parity["c_oracle"] = {
                "scans": sample,
                "labels_equal": bool(np.array_equal(label[:sample].cpu().numpy(), lab_c)),
                "known_equal": bool(np.array_equal(known[:sample].cpu().numpy().astype(bool), known_c)),
                "max_abs_dproba": float(np.abs(proba[:sample].cpu().numpy().astype(np.float64) - P_c).max())}
            if world > 1:
                r = world - 1
                other = device_cubes(sample, 1234 + r, dev).cpu().numpy()    # first scans of rank r's seeded shard
                _, lab_r, _, _, _ = c_oracle.scan_path(other, cm, mode="max")
                parity["remote_shard"] = {"rank": r, "scans": sample,
                                          "labels_equal": bool(np.array_equal(
                                              gathered[r * B:r * B + sample].cpu().numpy(), lab_r))}
                del other
        except Exception as e:      # the checker's checker must never take the bench line down
            parity["c_oracle"] = {"error": repr(e)}
            dtc = None
        if world == 1:
            cores = cpu_baseline.host_cores()
            v, dt, used = cpu_baseline.run_all_cores(host_np[:sample], cal, synth.CLASSES, cores=cores, repeats=3)
            v1, dt1, lat1 = cpu_baseline.run_one_core(host_np[:96], cal, synth.CLASSES, repeats=3)
            cpu = {"value": v, "unit": "scans/s", "cores": used, "kind": "port",
                   "cpu_model": cpu_baseline.cpu_model_string(),
                   "sample": "%d of the GPU-scored scans, per-scan predict.py loop (ndimage.zoom called like "
                             "common.py:143), one process per core, sklearn libsvm, best of 3 passes (%.1f s)" % (sample, dt),
                   "one_core": {"value": v1, "unit": "scans/s", "cores": 1, "ms_per_scan_median": lat1,
                                "sample": "96 scans, the reference-exact loop as predict.py runs it, best of 3 (%.1f s)" % dt1}}
            if dtc:
                cpu["c_port"] = {"value": sample / dtc, "unit": "scans/s", "cores": cores,
                                 "sample": "%d scans, oracle/c/radar_oracle.c, OpenMP over scans (%.1f s)" % (sample, dtc)}
            # ---- single-scan latency through the drop-in seam (configs[0]'s counterpart)
            from radar_ml_b200 import predict as rpredict
            from radar_ml_b200.model import GpuCalibratedClassifier
            gm = GpuCalibratedClassifier(params, engine=eng)
            le = synth.LabelEncoderLike()
            obs = synth.features(*synth.project_max(host_np[:1]))
            ijk1 = np.array([[11, 15, 88]], dtype=np.int32)

            def med_ms(fn, n):
                ts = []
                for _ in range(n):
                    t0 = time.perf_counter()
                    fn()
                    ts.append(time.perf_counter() - t0)
                return 1e3 * float(np.median(ts))
            rpredict.classifier(obs, gm, le)
            latency = {"unit": "ms, median", "classifier_B1": med_ms(lambda: rpredict.classifier(obs, gm, le), 1000),
                       "predict_targets_host_T1": med_ms(lambda: eng.predict_targets_host(host_np[0], ijk1), 500)}
            for nb_ in (1, 8, 64):
                o_ = (np.empty((nb_, C), np.float32), np.empty((nb_,), np.int32), np.empty((nb_,), np.uint8))
                eng.predict_host(host_np[:nb_], mode="max", out=o_)
                latency["predict_host_B%d" % nb_] = med_ms(lambda: eng.predict_host(host_np[:nb_], mode="max", out=o_), 200)
            name_g, p_g = rpredict.classifier(obs, gm, le)
            preds = cal.predict_proba(obs.reshape(1, -1))[0]
            latency["sklearn_classifier_B1"] = med_ms(lambda: cal.predict_proba(obs.reshape(1, -1)), 30)
            latency["classifier_matches_sklearn"] = bool(abs(float(p_g) - float(preds.max())) < 1e-5)
            latency["reference_loop_one_core"] = lat1
            eng.load_model(params)

    # ---- DerivedTarget axis sums (common.py:45-80) inside K1's single pass
    derived = None
    if rank == 0 and world == 1:
        from radar_ml_b200._lib import U8 as _U8
        nd = min(16384, B)
        cd = cubes[:nd]

        def ev_ms(fn, reps=5):
            fn()
            torch.cuda.synchronize(dev)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            for _ in range(reps):
                fn()
            a1.record(stream)
            torch.cuda.synchronize(dev)
            return a0.elapsed_time(a1) / reps
        t_proj = ev_ms(lambda: eng.project(cd, dtype=_U8))
        t_der = ev_ms(lambda: eng.derive_targets(cd, num_targets=3))
        t_one = ev_ms(lambda: eng.project_derive(cd, num_targets=3, dtype=_U8))
        _, _, ijk_f, sums_f = eng.project_derive(cd[:64], num_targets=3, dtype=_U8, want_sums=True)
        c64 = cd[:64].cpu().numpy().astype(np.float64)
        want = np.concatenate([c64.sum(axis=(2, 3)), c64.sum(axis=(1, 3)), c64.sum(axis=(1, 2))], axis=1)
        peak_d, _ = measured_peaks()
        derived = {"workload": "%d cubes: MAX projections (u8 rows) + DerivedTarget axis sums and top-3 indices" % nd,
                   "two_passes_ms": t_proj + t_der, "project_ms": t_proj, "derive_targets_ms": t_der,
                   "one_pass_ms": t_one, "one_pass_hbm_frac": nd * CUBE_BYTES / t_one / 1e6 / peak_d,
                   "sums_equal_numpy": bool(np.array_equal(sums_f.cpu().numpy().astype(np.float64), want)),
                   "ijk_equal_separate_kernel": bool(torch.equal(ijk_f, eng.derive_targets(cd[:64], num_targets=3)))}

    # ---- general-precision case: real-valued cubes + non-integral support vectors
    general = None
    if rank == 0 and world == 1:
        import copy
        p2 = copy.deepcopy(params)
        wob = 3e-5 * np.sin(np.arange(p2.sv.size, dtype=np.float64)).reshape(p2.sv.shape)
        p2.sv = np.clip(p2.sv + wob, 0.0, 1.0)                 # like augmented training data
        eng2 = Engine(local)
        eng2.load_model(p2)
        assert not eng2.model_is_integral
        Bg = int(os.environ.get("RML_BENCH_GENERAL", "65536"))
        cg = device_cubes(Bg, 777, dev, integer=False)
        outg = eng2.predict(cg)
        eng2.check_status()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        g0.record(stream)
        for _ in range(5):
            eng2.predict(cg, out=outg)
        g1.record(stream)
        torch.cuda.synchronize(dev)
        gms = g0.elapsed_time(g1) / 5
        po = restate.export_params(cal)
        po.sv = p2.sv
        ng = 96
        _, lab_o, _, _, P_o = restate.scan_path(cg[:ng].cpu().numpy(), po, mode="max")
        peak_g, _ = measured_peaks()
        general = {"workload": "%d real-valued cubes, non-integral support vectors: K1 f32 -> 24-bit "
                               "fixed-point digit planes -> exact multi-digit tcgen05 scorer" % Bg,
                   "value": Bg / gms * 1e3, "unit": "scans/s", "ms_per_step": gms,
                   "path_frac": Bg / gms * 1e3 * ALGO_BYTES_PER_SCAN / (peak_g * 1e9),
                   "labels_equal": bool(np.array_equal(outg[1][:ng].cpu().numpy(), lab_o)),
                   "max_abs_dproba": float(np.abs(outg[0][:ng].cpu().numpy().astype(np.float64) - P_o).max())}
        del cg
        eng2.close()

    # ---- uint8 cubes (the sensor's integers kept as bytes; include/radarml.h rml_predict*_u8):
    # same scans, a quarter of the bytes over PCIe and HBM, results identical bit for bit
    u8leg = None
    if rank == 0 and world == 1 and os.environ.get("RML_BENCH_U8", "1") != "0":
        host8 = torch.empty((Be, 22, 31, 176), dtype=torch.uint8, pin_memory=True)
        host8.copy_(cubes[:Be].to(torch.uint8))
        torch.cuda.synchronize(dev)
        host8_np = host8.numpy()
        out8 = (np.empty((Be, C), np.float32), np.empty((Be,), np.int32), np.empty((Be,), np.uint8))
        for _ in range(2):
            eng.predict_host(host8_np, mode="max", out=out8)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            eng.predict_host(host8_np, mode="max", out=out8)
        e2e8_s = time.perf_counter() - t0
        cubes8 = torch.empty((B, 22, 31, 176), device=dev, dtype=torch.uint8)
        for lo in range(0, B, 4096):
            cubes8[lo:lo + 4096] = cubes[lo:lo + 4096].to(torch.uint8)
        local_step()
        torch.cuda.synchronize(dev)
        lab_f32, proba_f32 = label.clone(), proba[:B].clone()
        res8 = eng.predict(cubes8)
        eng.check_status()
        for _ in range(2):
            eng.predict(cubes8, out=res8)
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        u0.record(stream)
        for _ in range(args.steps):
            eng.predict(cubes8, out=res8)
        u1.record(stream)
        torch.cuda.synchronize(dev)
        u8_ms = u0.elapsed_time(u1) / args.steps
        k1u = []
        for _ in range(5):
            eng.predict(cubes8, out=res8)
            rc = lib.rml_last_timing(ctx, Ct.byref(k1v), Ct.byref(totv), Ct.byref(fusedv))
            assert rc == 0, lib.rml_last_error(ctx)
            k1u.append(k1v.value)
        k1u_ms = float(np.mean(k1u))
        peak_u, _ = measured_peaks()
        u8leg = {"workload": "the same scans held as uint8 cubes (120 032 B/scan)",
                 "e2e": {"value": Be * args.steps / e2e8_s, "unit": "scans/s",
                         "h2d_bytes_per_step": Be * CUBE_BYTES // 4, "d2h_bytes_per_step": Be * (4 * C + 4 + 1)},
                 "value": B / u8_ms * 1e3, "unit": "scans/s", "ms_per_step": u8_ms,
                 "roofline": {"bound": "hbm", "kernel": "k1_project_max_u8in<u8>",
                              "achieved": B * (CUBE_BYTES // 4) / (k1u_ms * 1e-3) / 1e9, "peak": peak_u,
                              "unit": "GB/s", "frac": B * (CUBE_BYTES // 4) / (k1u_ms * 1e-3) / 1e9 / peak_u,
                              "k1_ms": k1u_ms, "fused_pipeline": bool(fusedv.value)},
                 "labels_equal_f32_path": bool(torch.equal(res8[1], lab_f32)),
                 "proba_equal_f32_path": bool(torch.equal(res8[0], proba_f32)),
                 "e2e_labels_equal": bool(np.array_equal(out8[1], e2e_labels))}
        del cubes8, host8

    # ---- config 4's per-GPU shard on one GPU: the like-for-like base of the scaling curve
    shard = None
    if world == 1 and B_alloc > B:
        lab_big = gathered[:B_alloc]
        for _ in range(2):
            local_step(B_alloc, lab_big)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        s0.record(stream)
        for _ in range(args.steps):
            local_step(B_alloc, lab_big)
        s1.record(stream)
        torch.cuda.synchronize(dev)
        sms_ = s0.elapsed_time(s1) / args.steps
        shard = {"workload": "configs[3]'s per-GPU shard (%d scans) on one GPU" % B_alloc,
                 "value": B_alloc / sms_ * 1e3, "unit": "scans/s", "ms_per_step": sms_,
                 "slice_checksum": str(label_checksum(lab_big))}

    # ---- the network configs: dnn at N = 1, sgan c_model at every N
    dnn_leg = sgan_leg = None
    if os.environ.get("RML_BENCH_NETS", "1") != "0":
        del work
        eng._work = None
        torch.cuda.empty_cache()
        n_res = cubes_all.shape[0]
        if world == 1:
            passes = max(1, 262144 // n_res)
            dnn_leg = net_leg("dnn", dev, cubes_all, n_res, passes, max(1, min(args.steps, 3)), rank, world, dist, stream)
        sgan_leg = net_leg("sgan_c", dev, cubes_all, min(16384, n_res), 1, max(1, min(args.steps, 5)), rank, world, dist, stream)

    if rank == 0:
        peak, peak_src = measured_peaks()
        k1_gbs = B * CUBE_BYTES / (k1_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": ("configs[1]: batch=65536 cubes 22x31x176 fp32, SVC-RBF 3-class, MAX projections"
                                    if world == 1 else
                                    "configs[3]: %d cubes/GPU sharded over %d GPUs, SVC-RBF, all-gather of labels" % (B, world)),
                       "scans_per_gpu": B, "global_batch": world * B, "n_sv": params.n_sv,
                       "features": params.n_features, "parallelism": "dp%d" % world,
                       "l2": "inputs (%.1f GB/GPU) larger than L2, no flush needed" % (B * CUBE_BYTES / 1e9),
                       "e2e_scans_per_step": Be,
                       "exchange": "rml_allgather_labels (ncclAllGather through the C ABI), labels written in place "
                                   "into the gather buffer" if world > 1 else "none (1 GPU)"},
            "roofline": {"bound": "hbm", "kernel": "k1_project_max<u8>", "achieved": k1_gbs, "peak": peak,
                         "unit": "GB/s", "frac": k1_gbs / peak, "traffic": B * K1_DRAM_TRAFFIC_PER_SCAN,
                         "traffic_unit": "bytes per launch (ncu dram read+write, profiles/r2_k1_project_max_u8_ncu.txt)",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_scan": CUBE_BYTES, "k1_ms": k1_ms,
                         "k2_exposed_ms": k2_ms, "fused_pipeline": bool(fused),
                         "path_frac": (value / world) * ALGO_BYTES_PER_SCAN / (peak * 1e9)},
            "e2e": {"value": e2e_value, "unit": "scans/s", "h2d_bytes_per_step": int(xfer["h2d_bytes"]),
                    "host_input_bytes_per_step": Be * CUBE_BYTES,
                    "d2h_bytes_per_step": Be * (4 * C + 4 + 1),
                    "host_narrowing": {"note": "rml_predict_host's default for float32 cubes on the integer path: host threads "
                                               "convert each 512-scan chunk to bytes (every value checked to be an integer in "
                                               "[0,255]; anything else goes over as float32) while the previous chunk is copied "
                                               "and scored, so a quarter of the bytes crosses PCIe; it switches itself off where "
                                               "the host converts slower than the bus moves float32 (rml_set_host_narrowing)",
                                       "active": xfer["active"], "threads": xfer["threads"],
                                       "narrowed_scans_per_step": int(xfer["narrowed_scans"]),
                                       "convert_gbs_of_float32_input": xfer["convert_gbs"],
                                       "results_equal_plain_copy": narrow_equal},
                    "plain_copy": {"note": "the same call with rml_set_host_narrowing(0): every float32 byte crosses the bus",
                                   "value": world * Be * args.steps / e2e_f32_s, "unit": "scans/s",
                                   "h2d_bytes_per_step": int(xfer_f["h2d_bytes"])},
                    "host_input_gbs_rank0": in_rank,
                    "h2d_gbs_per_rank_min": h2d_min, "h2d_gbs_sum": h2d_sum,
                    "host_copy_ceiling_gbs_per_rank_min": ceil_min, "host_copy_ceiling_gbs_sum": ceil_sum,
                    "ceiling_note": "the same pinned float32 buffers copied H2D by all ranks at once with no kernels: what the "
                                    "host memory / PCIe side of this box delivers; plain_copy can at best equal it, the "
                                    "narrowed default moves a quarter of the bytes (h2d_gbs_* = bytes on the bus, "
                                    "host_input_gbs_rank0 = float32 bytes consumed from host memory)",
                    "cpu_binding": binding},
            "gpu_launches": int(launches),
            "clocks": clocks, "parity": parity,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if latency:
            line["latency"] = latency
        if general:
            line["general_precision"] = general
        if derived:
            line["derived_targets"] = derived
        if u8leg:
            line["u8_cubes"] = u8leg
        if shard:
            line["shard_131072"] = shard
        if dnn_leg:
            line["dnn"] = dnn_leg
        if sgan_leg:
            line["sgan"] = sgan_leg
        emit(line)
    if world > 1:
        eng.lib.rml_comm_destroy(eng.ctx)
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scans-per-gpu", type=int, default=0)
    ap.add_argument("--e2e-scans", type=int, default=4096)
    ap.add_argument("--skip-extras", action="store_true",
                    help="profiling aid: skip the e2e, parity and cpu_baseline legs")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
