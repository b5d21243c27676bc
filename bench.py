#!/usr/bin/env python
"""bench.py -- image-pairs/sec of the weighted 8-point + pose path (F + R,t) on B200.

Contract (one JSON line on rank 0):
  python bench.py [--gpus N] [--steps K] [--warmup W]            ours (CUDA kernels via the C ABI)
  python bench.py --impl reference [...]                         the UNMODIFIED reference's CPU path (oracle/_ref; the
                                                                 oracle port only if that copy is absent)
  torchrun --nproc-per-node N bench.py --gpus N ...              one rank per GPU, weak scaling
  --workload C2 (default, the contract line) | C3 (64 x 2000, same step) | C4 (DeepFNet forward) | C5 (training step)

Timing: the K steps are recorded as ONE CUDA graph; that graph is replayed R times back to back so that the timed
region lasts >= 20 ms (`replays` in the line; `value` = pairs of K*R steps / time).  `value` issues the steps of a graph
on 8 parallel branches (independent batches); `value_serial` is the same measurement with ONE branch, i.e. launches
strictly back to back.  `bwd` is the forward + backward step (fit with saved state, loss head, head backward, fit
backward).  With more than one rank the line also carries `training_step`: the C5 step with its gradient all-reduce.

Workload (BASELINE.json configs[1]): one STEP = one batch of 256 image pairs x 1000 synthetic
correspondences (KITTI-shaped intrinsics, 0.5 px noise, 30 % outliers) through
  fepe_fit_fwd   (Hartley, Gram, eigenvector, rank 2, residual + epipolar residual)   and
  fepe_pose_fwd  (E = K^T F K, E -> R,t, pose errors vs GT, F-loss on 100 virtual points).
`value` is timed with the batches already resident in HBM: a ring of distinct batches larger than
the 126 MB L2 is cycled so no step finds its inputs in cache.  `e2e` times the same step through
the public host API with PINNED HOST buffers, H2D of the step's inputs and D2H of its results
inside the timed region.  `roofline` is for the dominant kernel at this launch size (the one-CTA-per-pair
latency kernel fepe_fit_fwd_small_kernel, pose head fused); `roofline_saturating` is the same entry point
at a batch that fills the machine, where it runs the split pipeline fepe_gram_kernel -> fepe_solve_kernel ->
fepe_resid_kernel (a 256-pair launch moves 7 MB, ~1 us of HBM time, and is latency bound by
nature).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np
import torch

METRIC = "image_pairs_per_sec_F_Rt"
UNIT = "pairs/s"
L2_BYTES = 126 * 1024 * 1024
CLAMP_EPI = 0.5      # in-model epipolar clamp (DeepFNet.py:479 default)
CLAMP_LOSS = 0.02    # configs/kitti_corr_baseline.yaml:36


_OUT = None


def emit(line: dict) -> None:
    out = _OUT if _OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=256, help="pairs per step (config C2: 256)")
    ap.add_argument("--ncorr", type=int, default=1000, help="correspondences per pair (C2: 1000)")
    ap.add_argument("--sat-batch", type=int, default=32768, help="batch of the saturating roofline run")
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying CUDA graphs")
    ap.add_argument("--streams", type=int, default=8,
                    help="parallel branches the K independent steps are issued on (1 = strictly back to back)")
    ap.add_argument("--no-extras", action="store_true", help="skip e2e / cpu_baseline / saturating legs")
    ap.add_argument("--phases", action="store_true", help="print per-phase SM cycles of the fused kernel")
    ap.add_argument("--workload", choices=["C2", "C3", "C4", "C5"], default="C2",
                    help="C2 (default, the contract line): solver only, 256 x 1000.  C3: the same step at 64 pairs x 2000 "
                         "correspondences (pose loss).  C4: whole DeepFNet forward (depth 5, "
                         "ErrorEstimator on tcgen05 + 5 fits), batch 512 x N=1000.  C5: training step (forward, F-loss, backward "
                         "through the analytic fit backward, one flattened NCCL all-reduce, Adam), 16 pairs per GPU.  "
                         "C3 / C4 / C5 are extra lines, not the contract")
    ap.add_argument("--min-ms", type=float, default=20.0, help="minimum duration of the timed region (graph replays)")
    ap.add_argument("--mlp", choices=["tc32", "bf16", "torch"], default="tc32",
                    help="C4 / C5: weight-MLP path (tc32 = split-fp16 tcgen05 at fp32 parity, the default of the model)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# stdout carries exactly one JSON line: NCCL's version banner / debug log (stdout by default) goes to stderr instead
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler(threading.Thread):
    """Polls NVML for SM clock / throttle reasons while the benchmark runs."""

    def __init__(self, index: int, period_s: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples = []          # (t, sm_mhz, reasons_bitmask, util)
        self.stop_flag = False
        self.sm_max = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                clk = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                self.samples.append((time.time(), clk, rs, util))
            except Exception:
                pass
            time.sleep(self.period)

    def summary(self, t0: float, t1: float) -> dict:
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "window": "unavailable"}
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        window = "timed_region"
        if len(inside) < 3:      # region shorter than the sampling period: use the loaded part of the run
            inside = [s for s in self.samples if s[3] > 0] or self.samples
            window = "whole_bench_under_load"
        names = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
                 0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
        mask = 0
        for s in inside:
            mask |= s[2]
        reasons = [n for bit, n in names.items() if mask & bit]
        return {"sm_mhz": statistics.median(s[1] for s in inside), "sm_max_mhz": self.sm_max,
                "reasons": reasons, "window": window, "samples": len(inside)}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy kernel)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(batch: int, ncorr: int):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if one matches."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return t.get(f"fit_fwd_B{batch}_N{ncorr}")
    except Exception:
        return None


def fit_bytes(B: int, N: int) -> int:
    # algorithmic bytes of one fused forward launch: 16 B coords + 4 B weight read, 4 B residual +
    # 4 B epipolar residual written per correspondence, 36 B F per pair (28 N + 36)
    return B * (28 * N + 36)


# ------------------------------------------------------------------------------------------------
def make_host_batches(n_batches: int, B: int, N: int, seed0: int):
    from fepe_b200 import synth
    return [synth.make_batch(B, N, seed=seed0 + i, weight_mode="softmax") for i in range(n_batches)]


class DeviceBatch:
    """One step's inputs resident on the device plus its caller-owned outputs."""

    def __init__(self, d: dict, dev):
        from fepe_b200 import _lib
        t = lambda k: torch.from_numpy(np.ascontiguousarray(d[k])).to(dev)
        self.m = t("matches_xy_ori")
        self.w = t("weights").reshape(self.m.shape[0], -1).contiguous()
        self.K, self.q, self.tt, self.Rt = t("Ks"), t("q_cam"), t("t_cam"), t("delta_Rtijs_4_4")
        self.v1, self.v2 = t("pts1_virt"), t("pts2_virt")
        B, N = self.m.shape[0], self.m.shape[1]
        self.F = torch.empty(B, 3, 3, device=dev)
        self.res = torch.empty(B, N, device=dev)
        self.epi = torch.empty(B, N, device=dev)
        self.pose = torch.empty(1, B, _lib.POSE_OUT_FLOATS, device=dev)
        self.saved = None               # backward state, allocated by enable_backward()

    def enable_backward(self, seed: int):
        """Buffers of the forward + backward step: the fit's saved state and fixed upstream gradients of the residual /
        epipolar-residual rows (in the model they come from the next layer's network)."""
        from fepe_b200 import _lib
        B, N = self.m.shape[0], self.m.shape[1]
        dev = self.m.device
        g = torch.Generator(device=dev).manual_seed(seed)
        self.saved = torch.empty(B, _lib.SAVED_DOUBLES, dtype=torch.float64, device=dev)
        self.gres = torch.randn(B, N, device=dev, generator=g) * 1e-3
        self.gepi = torch.randn(B, N, device=dev, generator=g) * 1e-3
        self.gq = torch.full((1, B), 1.0 / B, device=dev)
        self.gt = torch.full((1, B), 0.1 / B, device=dev)
        self.gl = torch.full((1, B), 1.0 / B, device=dev)
        self.dF = torch.empty(1, B, 3, 3, device=dev)
        self.gw = torch.empty(B, N, device=dev)


def step_fit(db: DeviceBatch, aff):
    from fepe_b200 import ops
    ops.fit_forward(db.m, db.w, aff, clamp_at=CLAMP_EPI, out=(db.F, db.res, db.epi, None))


def step_full(db: DeviceBatch, aff):
    from fepe_b200 import ops
    # one call of the C ABI (fepe_fit_pose_fwd); at the config batch it is ONE kernel (pose head fused into the fit)
    ops.fit_pose_forward(db.m, db.w, aff, db.K, db.q, db.tt, db.Rt, db.v1, db.v2, clamp_at=CLAMP_EPI,
                         virt_clamp_at=CLAMP_LOSS, out=(db.F, db.res, db.epi, None, db.pose[0]))


def step_fwd_bwd(db: DeviceBatch, aff):
    """Forward + backward of the solver step through the public ops: fit (saving its per-pair state) + loss head, then
    the head's backward (dL/dF of the q / t L2 errors and the F-loss) and the fit's backward (dL/dweights)."""
    from fepe_b200 import ops
    ops.fit_pose_forward(db.m, db.w, aff, db.K, db.q, db.tt, db.Rt, db.v1, db.v2, clamp_at=CLAMP_EPI,
                         virt_clamp_at=CLAMP_LOSS, out=(db.F, db.res, db.epi, db.saved, db.pose[0]))
    ops.pose_backward(db.F.unsqueeze(0), db.K, aff, db.q, db.tt, db.v1, db.v2, CLAMP_LOSS, db.pose, db.gq, db.gt, db.gl,
                      out=db.dF)
    ops.fit_backward(db.m, db.w, db.saved, db.dF[0], db.gres, db.gepi, aff, CLAMP_EPI, out=(db.gw, None))


def bwd_bytes(B: int, N: int) -> int:
    # forward 28 N + 36 (+ 512 B saved state written) and backward 32 N (20 B inputs + the two upstream rows read, 4 B
    # weight gradient written) + 512 B state read + 36 B dF per pair
    return B * (28 * N + 36 + 512) + B * (32 * N + 512 + 36)


def capture(fn):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g


def capture_pipelined(step_fns, n_steps: int, offset: int, n_branch: int):
    """ONE CUDA graph holding `n_steps` step launches: step i runs batch (offset + i) % len(step_fns) on branch
    i % n_branch.  The steps are independent batches, so the branches may overlap: the tail of one 256-CTA launch
    (its slowest pair) and the launch latency of the next no longer leave SMs idle."""
    main = torch.cuda.Stream()
    side = [torch.cuda.Stream() for _ in range(n_branch)]
    g = torch.cuda.CUDAGraph()
    n = len(step_fns)
    with torch.cuda.graph(g, stream=main):
        for st in side:
            st.wait_stream(main)
        for i in range(n_steps):
            with torch.cuda.stream(side[i % n_branch]):
                step_fns[(offset + i) % n]()
        for st in side:
            main.wait_stream(st)
    return g


def timed_loop(callables, steps: int, warmup: int, barrier=None):
    """CUDA-event time of `steps` calls cycling through `callables`; returns (seconds, t0_wall, t1_wall)."""
    n = len(callables)
    for i in range(warmup):
        callables[i % n]()
    torch.cuda.synchronize()
    if barrier is not None:
        barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.time()
    e0.record()
    for i in range(steps):
        callables[(warmup + i) % n]()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    if barrier is not None:
        barrier()
    return e0.elapsed_time(e1) * 1e-3, t0, t1


# ------------------------------------------------------------------------------------------------
def measure_steps(step_fns, K: int, W: int, n_branch: int, min_ms: float, barrier, reduce_max, no_graph: bool):
    """Time the step: W warm-up steps, then the K-step unit replayed R times back to back so that the timed region
    lasts >= min_ms (R is agreed over the ranks).  With graphs, len(step_fns) // K graphs tile the ring of batches, so
    consecutive replays keep walking through DISTINCT batches (> L2) instead of re-running the same K.
    Returns (seconds, steps timed = K * R, R, t0_wall, t1_wall)."""
    n = len(step_fns)
    if no_graph:
        one, _, _ = timed_loop(step_fns, K, W)
        R = int(reduce_max(max(1, int(np.ceil(1.3 * min_ms * 1e-3 / max(one, 1e-9))))))
        secs, t0, t1 = timed_loop(step_fns, K * R, 0, barrier)
        return secs, K * R, R, t0, t1
    G = 1 if K >= n else max(1, n // K)
    g_warm = capture_pipelined(step_fns, W, 0, n_branch)
    graphs = [capture_pipelined(step_fns, K, W + j * K, n_branch) for j in range(G)]
    torch.cuda.synchronize()
    g_warm.replay()
    one, _, _ = timed_loop([graphs[-1].replay], 1, 0)
    R = int(reduce_max(max(1, int(np.ceil(1.3 * min_ms * 1e-3 / max(one, 1e-9))))))
    secs, t0, t1 = timed_loop([g.replay for g in graphs], R, 0, barrier)
    return secs, K * R, R, t0, t1


def workload_shape(args):
    """(B, N, label) of the solver workloads: C2 = BASELINE.json configs[1], C3 = configs[2]."""
    if args.workload == "C3":
        B = 64 if args.batch == 256 else args.batch
        N = 2000 if args.ncorr == 1000 else args.ncorr
        return B, N, f"C3: batch={B} pairs x N={N} corr, 30% outliers, Fit + epi residual + F-loss + E->R,t (pose loss)"
    return args.batch, args.ncorr, (f"C2: batch={args.batch} pairs x N={args.ncorr} corr, 30% outliers, "
                                    "Fit + epi residual + F-loss + E->R,t")


def bench_config(label: str, B: int, N: int) -> dict:
    """The keys both arms share (the driver compares the two `config` dicts)."""
    return {"workload": label, "batch_per_gpu": B, "ncorr": N}


# ------------------------------------------------------------------------------------------------
_REF = {"tried": False, "ns": None, "fit": None, "nhw": None}


def reference_modules():
    """The UNMODIFIED reference (oracle/_ref, made by oracle/make_ref.py) or None."""
    if not _REF["tried"]:
        _REF["tried"] = True
        try:
            from oracle import ref_env
            with ref_env.quiet():
                ns = ref_env.import_reference()
                _REF["fit"] = ns.Fit(is_cuda=False, if_cpu_svd=False)
            _REF["ns"] = ns
        except Exception as e:                                     # noqa: BLE001
            print(f"bench.py: unmodified reference unavailable ({e}); using the oracle port", file=sys.stderr)
    return _REF["ns"]


def reference_step(d: dict):
    """The reference's own CPU path for one batch: NormalizeAndExpand_HW + Fit(is_cuda=False) (DeepFNet.py:93-257) +
    compute_epi_residual (utils_F.py:400-413) + get_all_loss_DeepF (F-loss, E; train_good_utils.py:298) +
    get_Rt_loss(device='cpu') (:64), all UNMODIFIED from oracle/_ref.  Falls back to the oracle port only when that
    copy is absent.  Returns nothing; timed by the caller."""
    T = torch.from_numpy
    ns = reference_modules()
    if ns is None:
        from oracle import fepe_oracle as O
        with torch.no_grad():
            p1, p2, Tn = O.norm_hw(T(d["matches_xy_ori"]), d["image_size"])
            Fo, res = O.fit_weighted_svd(p1, p2, T(d["weights"]))
            O.epi_residual(p1, p2, Fo, CLAMP_EPI)
            _, _, E_layers = O.f_loss_layers([Fo], Tn, Tn, T(d["pts1_virt"]), T(d["pts2_virt"]), T(d["Ks"]), CLAMP_LOSS)
            O.pose_errors(E_layers[0], T(d["q_cam"]), T(d["t_cam"]), T(d["delta_Rtijs_4_4"]))
        return
    from oracle import ref_env
    with torch.no_grad(), ref_env.quiet():
        m = T(d["matches_xy_ori"])
        nhw = ns.NormalizeAndExpand_HW(d["image_size"], is_cuda=False)
        p1, p2, T1, T2 = nhw(m)
        p1, p2 = p1.permute(0, 2, 1).contiguous(), p2.permute(0, 2, 1).contiguous()
        w = T(d["weights"])
        Fo, res = _REF["fit"](p1, p2, w)
        epi = ns.utils_F.compute_epi_residual(p1, p2, Fo)
        outs = {"weights": w, "F_est": Fo, "T1": T1, "T2": T2, "out_layers": [Fo], "residual_layers": [res],
                "weights_layers": [w], "epi_res_layers": [epi.unsqueeze(1)]}
        lp = {"depth": 1, "clamp_at": CLAMP_LOSS, "if_tri_depth": False, "if_sample_loss": False}
        _, _, _, _, _, _, E_layers = ns.tgu.get_all_loss_DeepF(outs, T(d["pts1_virt"]), T(d["pts2_virt"]), T(d["Ks"]), lp,
                                                               get_residual_summaries=False)
        ns.tgu.get_Rt_loss(E_layers, T(d["Ks"]), m[:, :, :2], m[:, :, 2:], T(d["delta_Rtijs_4_4"]), T(d["q_cam"]),
                           T(d["t_cam"]), device="cpu")


def reference_kind():
    return ("reference", "the unmodified reference from oracle/_ref: NormalizeAndExpand_HW + Fit(is_cuda=False) + "
            "compute_epi_residual + get_all_loss_DeepF + get_Rt_loss(device='cpu')") if reference_modules() is not None \
        else ("port", "oracle/fepe_oracle.py (per-pair torch.svd loop, deepFEPE/models/DeepFNet.py:232-240)")


def pick_cpu_threads(d: dict) -> int:
    """torch.svd on 1000x9 matrices does not scale with threads (on a 128-core host the default is ~20x
    SLOWER than 8 threads): give the reference its best thread count among a few candidates."""
    cores = os.cpu_count() or 1
    probe = slice_batch(d, min(16, d["matches_xy_ori"].shape[0]))
    best_n, best_t = 1, float("inf")
    for n in sorted({1, 4, 8, 16, 32, cores}):
        if n > cores:
            continue
        torch.set_num_threads(n)
        reference_step(probe)
        t = time.perf_counter()
        reference_step(probe)
        dt = time.perf_counter() - t
        if dt < best_t:
            best_n, best_t = n, dt
    torch.set_num_threads(best_n)
    return best_n


def slice_batch(d: dict, n: int) -> dict:
    out = {}
    for k, v in d.items():
        out[k] = v[:n] if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == d["matches_xy_ori"].shape[0] else v
    return out


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    B, N, label = workload_shape(args)
    base = make_host_batches(2, B, N, seed0=1000)
    pick_cpu_threads(base[0])
    kind, how = reference_kind()
    # calibrate the per-step sample so that the whole run stays within ~2 minutes
    t = time.perf_counter()
    reference_step(slice_batch(base[0], min(B, 32)))
    per_pair = (time.perf_counter() - t) / min(B, 32)
    budget = 120.0 / max(1, args.steps + args.warmup)
    n_s = int(max(8, min(B, budget / max(per_pair, 1e-9))))
    sample = [slice_batch(b, n_s) for b in base]
    for i in range(args.warmup):
        reference_step(sample[i % 2])
    t0 = time.perf_counter()
    for i in range(args.steps):
        reference_step(sample[i % 2])
    dt = time.perf_counter() - t0
    value = n_s * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(label, B, N),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                         "host_cores": os.cpu_count(),
                         "sample": f"{n_s} of the {B} pairs of a step, {args.steps} steps; {how} on host cores"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        barrier = lambda: (dist.barrier(), torch.cuda.synchronize())
    else:
        barrier = None

    import __graft_entry__ as entry
    entry.build()
    from fepe_b200 import ops, synth, _lib

    B, N, label = workload_shape(args)
    K = args.steps
    aff = ops.hw_affine(synth.KITTI_IMAGE_SIZE)
    per_batch_bytes = B * N * 20
    ring_n = max(2, int(np.ceil(1.6 * L2_BYTES / per_batch_bytes)))
    ring_n = min(ring_n, 96)
    n_branch = 1 if args.no_graph else max(1, args.streams)
    if K < ring_n:
        ring_n = (ring_n + K - 1) // K * K                        # ring_n // K graphs tile the ring exactly
    else:
        ring_n = (ring_n + n_branch - 1) // n_branch * n_branch   # a batch's output buffers stay on one branch
    sampler = ClockSampler(local)
    sampler.start()

    def reduce_max(v: float) -> float:
        if not use_dist:
            return v
        tt = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    host = make_host_batches(ring_n, B, N, seed0=10_000 * (rank + 1))
    ring = [DeviceBatch(d, dev) for d in host]
    for db in ring[:2]:
        step_full(db, aff)          # first-use configuration happens outside any graph capture
    torch.cuda.synchronize()

    step_fns = [(lambda db=db: step_full(db, aff)) for db in ring]
    if args.no_graph:
        fit_calls = step_fns
    else:
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            fit_graphs = [capture(fn) for fn in step_fns]     # the timed step's own launch
        torch.cuda.synchronize()
        fit_calls = [g.replay for g in fit_graphs]

    # ---- the contract number: K-step unit, device timed, >= min_ms, max over ranks ---------------
    W = max(args.warmup, 3)
    secs, n_timed, replays, t0, t1 = measure_steps(step_fns, K, W, n_branch, args.min_ms, barrier, reduce_max,
                                                   args.no_graph)
    secs = reduce_max(secs)
    value = world * B * n_timed / secs
    # the same with ONE branch: launches strictly back to back (what a caller without parallel streams gets)
    if n_branch > 1:
        secs1, n1, _, _, _ = measure_steps(step_fns, K, W, 1, args.min_ms, barrier, reduce_max, args.no_graph)
        value_serial = world * B * n1 / reduce_max(secs1)
    else:
        value_serial = value

    # ---- dominant kernel alone at this launch size ----------------------------------------------
    n_alone = max(K, 200)
    fit_secs, _, _ = timed_loop(fit_calls, n_alone, 10)
    fit_us = fit_secs / n_alone * 1e6
    peak, peak_src = measured_peak_gbs()
    achieved = fit_bytes(B, N) / (fit_us * 1e-6) / 1e9
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    small = B <= (6 if N * 20 <= 28 * 1024 else 2) * sms and N * 20 <= 56 * 1024      # fepe_fit.cu: fit_fwd_impl
    fit_kernel = "fepe_fit_fwd_small_kernel<POSE>" if small else "fepe_fit_fwd_kernel + fepe_pose_fwd_kernel"
    traffic = ncu_traffic(B, N)
    roofline = {"bound": "hbm", "kernel": fit_kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "traffic_source": None if traffic is None else
                "static: profiles/traffic.json (dram__bytes of the same launch from the committed ncu --set full capture; "
                "not re-measured in this run)",
                "launch_us": fit_us,
                "algorithmic_bytes_per_launch": fit_bytes(B, N), "peak_source": peak_src,
                "note": "ONE launch of the timed step alone, single stream (batches of <= 6 pairs per SM go to the "
                        "one-CTA-per-pair latency kernel with the pose head fused into it, fepe_fit_pose_fwd); "
                        "roofline_saturating is the same entry point at a batch that fills the 148 SMs many times over"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": secs / n_timed * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 (Gram / eigen / SVD in f64)", "data": "synthetic",
        "config": bench_config(label, B, N),
        "replays": replays, "timed_steps": n_timed, "timed_region_ms": secs * 1e3,
        "value_serial": value_serial,
        "run": {"parallelism": f"pairs sharded over {world} GPU(s), no collective",
                "l2_policy": f"ring of {ring_n} distinct batches ({ring_n * per_batch_bytes / 2**20:.0f} MiB) > 126 MiB L2, "
                             "walked across the graph replays",
                "launch": "eager" if args.no_graph else
                f"value: CUDA graphs of K step launches on {n_branch} parallel branches (independent batches), replayed "
                f"{replays}x; value_serial: the same on 1 branch"},
        "gpu_launches": n_timed * (1 if small else 2),
        "roofline": roofline,
        # the same kernel inside the timed region: launches of independent batches overlap on the parallel branches, so
        # the AVERAGE time per launch (region / launches) is shorter than one launch alone
        "roofline_timed_region": {"bound": "hbm", "kernel": fit_kernel, "unit": "GB/s", "peak": peak,
                                  "achieved": fit_bytes(B, N) * n_timed / secs / 1e9,
                                  "frac": fit_bytes(B, N) * n_timed / secs / 1e9 / peak,
                                  "avg_launch_us": secs / n_timed * 1e6,
                                  "note": "algorithmic bytes of all launches of the timed region / its CUDA-event time "
                                          f"({n_branch} concurrent branches on this rank)"},
    }

    if not args.no_extras:
        # ---- forward + backward of the same step ("fwd and fwd+bwd") --------------------
        for i, db in enumerate(ring):
            db.enable_backward(777 + i)
        step_fwd_bwd(ring[0], aff)
        torch.cuda.synchronize()
        bwd_fns = [(lambda db=db: step_fwd_bwd(db, aff)) for db in ring]
        bsecs, bn, brep, _, _ = measure_steps(bwd_fns, K, W, n_branch, args.min_ms, barrier, reduce_max, args.no_graph)
        bsecs = reduce_max(bsecs)
        if args.no_graph:
            b_alone = bwd_fns
        else:
            with torch.cuda.stream(torch.cuda.Stream()):
                b_graphs = [capture(fn) for fn in bwd_fns[:max(8, min(len(bwd_fns), 48))]]
            torch.cuda.synchronize()
            b_alone = [g.replay for g in b_graphs]
        b_secs1, _, _ = timed_loop(b_alone, 200, 10)
        b_us = b_secs1 / 200 * 1e6
        b_ach = bwd_bytes(B, N) / (b_us * 1e-6) / 1e9
        line["bwd"] = {"value": world * B * bn / bsecs, "unit": UNIT, "ms_per_step": bsecs / bn * 1e3, "replays": brep,
                       "what": "forward (fit saving its state + fused loss head) + fepe_pose_bwd + fepe_fit_bwd, same "
                               "batches, same graph / branch structure as `value`",
                       "kernels": ("fepe_fit_fwd_small_kernel<POSE>" if small else "fepe_fit_fwd_kernel + fepe_pose_fwd_kernel")
                       + " + fepe_pose_bwd_kernel + fepe_fit_bwd_kernel<0>",
                       "gpu_launches_per_step": (1 if small else 2) + 2,
                       "roofline": {"bound": "hbm", "achieved": b_ach, "peak": peak, "unit": "GB/s", "frac": b_ach / peak,
                                    "launch_us": b_us, "algorithmic_bytes_per_step": bwd_bytes(B, N),
                                    "note": "one forward + backward step alone on one stream; bytes: forward 28 N + 548, "
                                            "backward 32 N + 548 per pair"}}

    if not args.no_extras:
        # ---- end to end through the host API: pinned host -> device -> results on host ------------
        from fepe_b200.staging import StagedStep
        nbuf = 4
        V = host[0]["pts1_virt"].shape[1]
        # nbuf staging slots, each with its own pinned host buffer holding a DISTINCT batch (what a DataLoader with
        # pin_memory hands over) and its own stream; one step = one replay of the slot's captured
        # H2D -> fepe_fit_pose_fwd -> D2H graph (StagedStep.capture / replay)
        stages = [StagedStep(B, N, V, dev) for _ in range(nbuf)]
        streams = [torch.cuda.Stream() for _ in range(nbuf)]
        for j in range(nbuf):
            stages[j].pack(host[j % len(host)], out=stages[j].h_in)
            if not args.no_graph:
                stages[j].capture(streams[j], aff, CLAMP_EPI, CLAMP_LOSS)
        h2d, d2h = stages[0].in_bytes, stages[0].out_bytes

        def e2e_step(i):
            j = i % nbuf
            if args.no_graph:
                stages[j].run(streams[j], aff, CLAMP_EPI, CLAMP_LOSS)
            else:
                stages[j].replay()

        e2e_steps = max(200, min(args.steps, 400))
        # the PCIe ceiling of THIS box at this moment: the same pinned buffers copied with nothing else, issued exactly like
        # the e2e steps (one captured graph per slot, round-robin over the slots' streams)
        copy_graphs = []
        for j in range(nbuf):
            gc = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gc, stream=streams[j]):
                stages[j].d_in.copy_(stages[j].h_in, non_blocking=True)
            copy_graphs.append(gc)
        def copy_step(i):
            with torch.cuda.stream(streams[i % nbuf]):
                copy_graphs[i % nbuf].replay()

        for i in range(2 * nbuf):
            copy_step(i)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_copy = max(50, min(args.steps, 400))
        tc0 = time.perf_counter()
        c0.record()
        for i in range(n_copy):
            copy_step(i)
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)
        c1.record()
        torch.cuda.synchronize()
        h2d_only_us = max(c0.elapsed_time(c1) * 1e-3, time.perf_counter() - tc0) / n_copy * 1e6
        for i in range(6):
            e2e_step(i)
        torch.cuda.synchronize()
        if barrier is not None:
            barrier()
        tw0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(e2e_steps):
            e2e_step(i)
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)
        e1.record()
        torch.cuda.synchronize()
        e2e_secs = max(e0.elapsed_time(e1) * 1e-3, time.perf_counter() - tw0)
        e2e_secs = reduce_max(e2e_secs)
        line["e2e"] = {"value": world * B * e2e_steps / e2e_secs, "unit": UNIT, "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": e2e_secs / e2e_steps * 1e3,
                       "h2d_only_ms_per_step": h2d_only_us * 1e-3,
                       "n_gpus": world, "how": f"fepe_b200.staging.StagedStep: {nbuf} slots / streams, per step one pinned "
                                               "H2D copy + fepe_fit_pose_fwd + one D2H copy"
                                               + ("" if args.no_graph else ", replayed as one CUDA graph per slot")}


    if rank == 0 and world == 1 and not args.no_extras:
        # ---- saturating batch for the same kernel (roofline of the kernel itself) ---------------
        SB = args.sat_batch
        reps = (SB + B * ring_n - 1) // (B * ring_n)
        big_m = torch.cat([db.m for db in ring] * reps)[:SB].contiguous()
        big_w = torch.cat([db.w for db in ring] * reps)[:SB].contiguous()
        outF, outr, oute = torch.empty(SB, 3, 3, device=dev), torch.empty(SB, N, device=dev), torch.empty(SB, N, device=dev)
        sat_call = lambda: ops.fit_forward(big_m, big_w, aff, clamp_at=CLAMP_EPI, out=(outF, outr, oute, None))
        sat_secs, _, _ = timed_loop([sat_call], 10, 3)
        sat_us = sat_secs / 10 * 1e6
        sat_ach = fit_bytes(SB, N) / (sat_us * 1e-6) / 1e9
        line["roofline_saturating"] = {
            "bound": "hbm", "kernel": "fepe_gram_kernel + fepe_solve_kernel + fepe_resid_kernel (split pipeline of "
                                      "fepe_fit_fwd; launch_us is the three launches together)", "batch": SB, "achieved": sat_ach, "peak": peak,
            "unit": "GB/s", "frac": sat_ach / peak, "traffic": ncu_traffic(SB, N), "launch_us": sat_us,
            "pairs_per_sec": SB / (sat_us * 1e-6),
            "l2_policy": f"input {SB * N * 20 / 2**20:.0f} MiB per launch > L2"}
        if args.phases:
            sv = torch.empty(SB, _lib.SAVED_DOUBLES, dtype=torch.float64, device=dev)
            ops.fit_forward(big_m, big_w, aff, clamp_at=CLAMP_EPI, out=(outF, outr, oute, sv))
            torch.cuda.synchronize()
            ph = sv[:, 56:61].mean(0).tolist()
            # split pipeline: cycles of a pair-team inside fepe_gram_kernel (solve / residual are separate kernels)
            line["phase_cycles_saturating"] = dict(zip(["wait", "hartley", "gram", "reduce_store"], ph[:4]))
            line["factorisations_mean"] = float(sv[:, 52].mean())
        del big_m, big_w, outF, outr, oute

        # ---- the reference's CPU algorithm on this box's host cores (bounded sample) ---------------
        pick_cpu_threads(host[0])
        sample = [slice_batch(host[0], min(B, 128)), slice_batch(host[1], min(B, 128))]
        reference_step(sample[0])
        n_done, tc0 = 0, time.perf_counter()
        while time.perf_counter() - tc0 < 10.0 or n_done < 3:
            reference_step(sample[n_done % 2])
            n_done += 1
        cpu_secs = time.perf_counter() - tc0
        kind, how = reference_kind()
        line["cpu_baseline"] = {"value": sample[0]["matches_xy_ori"].shape[0] * n_done / cpu_secs, "unit": UNIT,
                                "cores": torch.get_num_threads(), "kind": kind, "host_cores": os.cpu_count(),
                                "sample": f"{n_done} x {sample[0]['matches_xy_ori'].shape[0]} pairs x N={N} of the same "
                                          f"workload through {how} in {cpu_secs:.1f} s"}
    if "e2e" not in line:
        line["e2e"] = None
    if use_dist and not args.no_extras:
        # the one collective north_star names: the training step (C5, 16 pairs per GPU) with its gradient all-reduce
        del ring
        torch.cuda.empty_cache()
        line["training_step"] = c5_measure(args, dev, rank, world, steps=10, warm=3, mlp=args.mlp)

    sampler.stop_flag = True
    sampler.join(timeout=1.0)
    line["clocks"] = sampler.summary(t0, t1)
    if use_dist:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(line)


def run_c4(args):
    """Config 4 of BASELINE.json: ErrorEstimator(4) + 4 x ErrorEstimator(7) + 5 fits + 4 epipolar residuals
    (DeepFNet.forward, depth 5) on a batch of 512 pairs x 1000 correspondences, inference mode."""
    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import __graft_entry__ as entry
    entry.build()
    from fepe_b200 import synth
    from fepe_b200.models import DeepFNet
    B, N = (512 if args.batch == 256 else args.batch), args.ncorr
    torch.manual_seed(0)
    net = DeepFNet(depth=5, image_size=list(synth.KITTI_IMAGE_SIZE), if_quality=False).cuda().eval()
    host = make_host_batches(2, min(B, 128), N, seed0=77)
    batches = [{"matches_xy_ori": torch.from_numpy(d["matches_xy_ori"]).to(dev).repeat((B + 127) // 128, 1, 1)[:B].contiguous()}
               for d in host]
    steps, warm = min(args.steps, 50), max(3, min(args.warmup, 10))
    sampler = ClockSampler(local)
    sampler.start()
    res = {}
    with torch.no_grad():
        for path in dict.fromkeys([args.mlp, "tc32", "bf16"]):
            net.set_mlp_path(path)
            secs, t0, t1 = timed_loop([(lambda b=b: net(b)) for b in batches], steps, warm)
            res[path] = (secs, t0, t1)
        net.set_mlp_path("torch")
        n32 = max(3, steps // 10)
        secs32, _, _ = timed_loop([(lambda b=b: net(b)) for b in batches], n32, 2)
    secs, t0, t1 = res[args.mlp]
    sampler.stop_flag = True
    sampler.join(timeout=1.0)
    flop_per_pt = lambda cin: 2 * (cin * 64 + 64 * 128 + 128 * 1024 + 1024 * 512 + 512 * 256 + 256)
    flops = B * N * (flop_per_pt(4) + 4 * flop_per_pt(7))
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak_tf = float(json.load(f)["bf16_tflops_sustained"])
    except Exception:
        peak_tf = 1400.0
    mma_per_flop = {"tc32": 3.0, "bf16": 1.0, "torch": 0.0}[args.mlp]
    dtype = {"tc32": "MLP: split-fp16 (hi + lo) operands, 3 x tcgen05 kind::f16 MMAs per product, fp32 accumulate, fp32 "
                     "activations, fp64 statistics = fp32 parity; solver f32 / f64",
             "bf16": "bf16 MLP (fp32 accumulate) + f32/f64 solver", "torch": "fp32 library MLP + f32/f64 solver"}[args.mlp]
    ach = flops / (secs / steps) / 1e12
    line = {"metric": METRIC, "value": world * B * steps / secs, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warm, "ms_per_step": secs / steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": {"workload": f"C4: DeepFNet forward depth 5 (5 ErrorEstimator evaluations on tcgen05 + 5 fused fits), "
                                   f"batch={B} x N={N}, inference", "batch_per_gpu": B, "ncorr": N},
            "mlp_path": args.mlp,
            "roofline": {"bound": "tensor", "kernel": "fepe_mlp32_gemm_kernel (whole step counted)" if args.mlp == "tc32"
                         else "fepe_mlp_gemm_persist_kernel (whole step counted)",
                         "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": None,
                         "executed_tensor_tflops": ach * mma_per_flop, "executed_frac": ach * mma_per_flop / peak_tf,
                         "note": "achieved = ALGORITHMIC MLP flops of the step (1.59 MFLOP per correspondence and "
                                 "evaluation) / step time, against the measured bf16 cuBLAS rate; the fp32-parity path "
                                 "executes 3 fp16 MMAs per product (executed_*), so its ceiling on this scale is 1/3; "
                                 "the step also contains the first / last layers, the statistics and the 5 fits"},
            "pairs_per_sec_by_mlp_path": {k: world * B * steps / v[0] for k, v in res.items()} |
                                         {"torch": world * B * n32 / secs32},
            # per ErrorEstimator evaluation (tc32): first, 5 x scale_shift, 4 x gemm, last = 11 launches; one fit launch
            # per DeepFNet iteration
            "gpu_launches": steps * (5 * 11 + 5),
            "clocks": sampler.summary(t0, t1)}
    if rank == 0:
        emit(line)


def c5_measure(args, dev, rank, world, steps: int, warm: int, mlp: str) -> dict:
    """Config 5 of BASELINE.json: end-to-end training step (random keypoints / descriptors stand in for SuperPoint),
    16 pairs per GPU (batch 128 over 8 GPUs); the only collective is ONE all-reduce of the flat gradient buffer
    (fepe_b200.dist.FlatGradients; the reference: nn.DataParallel, deepFEPE/train_good.py:309-314).  The process group
    must already exist when world > 1."""
    import torch.distributed as dist
    from fepe_b200 import synth, dist as fdist
    from fepe_b200.models import DeepFNet
    from fepe_b200.matching import get_matches_from_descriptors
    from fepe_b200.losses import deepf_training_loss
    from fepe_b200 import ops as fops
    B, N = 16, args.ncorr
    NKP, DESC = 1200, 256                    # keypoints per image and descriptor size (SuperPoint: 256)
    torch.manual_seed(0)
    # with_quality (configs/kitti_corr_baseline.yaml): the match score is the one quality channel
    net = DeepFNet(depth=5, image_size=list(synth.KITTI_IMAGE_SIZE), if_quality=True, quality_size=1).cuda()
    net.set_mlp_path(mlp)
    # configs/kitti_corr_baseline.yaml:62; fused=True: torch's own single-kernel Adam over all 44 parameters
    # (capturable: its step can be replayed as a CUDA graph as well, see below)
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, fused=True, capturable=not args.no_graph)
    adam_graph = [None]
    # gradients of all parameters are views of ONE buffer, and the MLP's weight-gradient kernels add straight into them
    flat = fdist.FlatGradients(net.parameters(), fuse_accumulation=True)
    aff = fops.hw_affine(synth.KITTI_IMAGE_SIZE)
    # "SuperPoint frozen / random desc": a synthetic two-view scene gives NKP corresponding keypoints; image 2's are
    # shuffled and carry noisy copies of image 1's random unit descriptors.  The step starts from keypoints + descriptors.
    d = synth.make_batch(B, NKP, seed=500 + rank)
    T = lambda k: torch.from_numpy(d[k]).to(dev)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    m_all = T("matches_xy_ori")
    perm = torch.stack([torch.randperm(NKP, device=dev, generator=gen) for _ in range(B)])
    kp1 = m_all[:, :, :2].contiguous()
    kp2 = torch.gather(m_all[:, :, 2:], 1, perm.unsqueeze(-1).expand(-1, -1, 2)).contiguous()
    desc1 = torch.nn.functional.normalize(torch.randn(B, NKP, DESC, device=dev, generator=gen), dim=2)
    desc2 = torch.gather(desc1, 1, perm.unsqueeze(-1).expand(-1, -1, DESC))
    desc2 = torch.nn.functional.normalize(desc2 + 0.03 * torch.randn(B, NKP, DESC, device=dev, generator=gen), dim=2).contiguous()
    v1, v2 = T("pts1_virt"), T("pts2_virt")
    Ks, Rt, q_cam, t_cam = T("Ks"), T("delta_Rtijs_4_4"), T("q_cam"), T("t_cam")

    n_matches = []

    def fwd_bwd(xs, quality):
        """Forward + losses + backward of one batch; the gradients accumulate into the flat buffer."""
        outs = net({"matches_xy_ori": xs, "quality": quality})
        # get_all_loss_DeepF (F-loss on the virtual points, E_i = K^T T2^T F_i T1 K: train_good_utils.py:325-364) +
        # get_Rt_loss (:64-295) + their combination (Train_model_pipeline.py:580-592, if_qt_loss): one launch of the
        # fused head over all (layer, pair) items, one in the backward
        loss, parts = deepf_training_loss(outs["out_layers"], Ks, v1, v2, Rt, q_cam, t_cam, aff, clamp_at=CLAMP_LOSS)
        mid = torch.cuda.Event(enable_timing=True) if not capturing[0] else None
        if mid is not None:
            mid.record()
        loss.backward()
        return loss.detach(), parts["R_angle"], parts["t_angle"], mid

    capturing = [False]
    graphed = [None]

    def step():
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        flat.zero_()
        ev[0].record()
        # train_good_utils.py:649-724: mutual-NN matches -> [B,N,4] + quality (fepe_nn_match, on the device)
        mt = get_matches_from_descriptors(kp1, kp2, desc1, desc2, 1.0, out_num_points=N, generator=gen)
        n_matches.append(mt["num_matches"])
        ev[1].record()
        if graphed[0] is not None:
            loss, r_ang, t_ang, _ = graphed[0].replay(mt["xs"], mt["quality"])
            ev[2] = None                                      # forward / backward are one graph launch
        else:
            loss, r_ang, t_ang, mid = fwd_bwd(mt["xs"], mt["quality"])
            ev[2] = mid
        metrics = (r_ang.cpu(), t_ang.cpu())                  # the angular metrics the reference logs every step (one D2H each)
        ev[3].record()
        flat.allreduce_mean_()
        ev[4].record()
        if adam_graph[0] is not None:
            adam_graph[0].replay()
        else:
            opt.step()
        ev[5].record()
        return ev

    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    launch = "eager"
    if not args.no_graph:
        # the forward + losses + backward of the step as ONE CUDA graph (the step is launch-bound at 16 pairs per GPU)
        try:
            from fepe_b200.graphs import GraphedStep
            mt0 = get_matches_from_descriptors(kp1, kp2, desc1, desc2, 1.0, out_num_points=N, generator=gen)
            flat.zero_()
            capturing[0] = True
            graphed[0] = GraphedStep(fwd_bwd, (mt0["xs"], mt0["quality"]))
            launch = "forward + losses + backward replayed as one CUDA graph (fepe_b200.graphs.GraphedStep); matcher, all-reduce eager"
        except Exception as e:                                 # noqa: BLE001
            graphed[0] = None
            launch = f"eager (graph capture failed: {type(e).__name__}: {str(e)[:120]})"
        capturing[0] = False
        if graphed[0] is not None:
            # torch's own fused Adam step as a second graph (its state exists: the warm-up steps ran it eagerly)
            try:
                torch.cuda.synchronize()
                ga = torch.cuda.CUDAGraph()
                with torch.cuda.graph(ga):
                    opt.step()
                adam_graph[0] = ga
                launch += "; Adam (torch, fused + capturable) replayed as a second graph"
            except Exception as e:                             # noqa: BLE001
                adam_graph[0] = None
                launch += f"; Adam eager ({type(e).__name__})"
        for _ in range(2):
            step()
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    evs = [step() for _ in range(steps)]
    e1.record()
    torch.cuda.synchronize()
    secs = fdist.max_over_ranks(e0.elapsed_time(e1) * 1e-3, dev)
    br = {"match": sum(ev[0].elapsed_time(ev[1]) for ev in evs) / steps,
          "allreduce": sum(ev[3].elapsed_time(ev[4]) for ev in evs) / steps,
          "adam": sum(ev[4].elapsed_time(ev[5]) for ev in evs) / steps}
    if evs[0][2] is not None:
        br["fwd"] = sum(ev[1].elapsed_time(ev[2]) for ev in evs) / steps
        br["bwd"] = sum(ev[2].elapsed_time(ev[3]) for ev in evs) / steps
    else:
        br["fwd_bwd_graph"] = sum(ev[1].elapsed_time(ev[3]) for ev in evs) / steps
    # the collective alone, back to back (device time, max over ranks): the step's own interval also contains the wait
    # for the slowest rank's backward
    ar_us = None
    gbytes = flat.flat.numel() * 4
    if world > 1:
        for _ in range(3):
            dist.all_reduce(flat.flat)
        torch.cuda.synchronize()
        dist.barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(20):
            dist.all_reduce(flat.flat)
        a1.record()
        torch.cuda.synchronize()
        ar_us = fdist.max_over_ranks(a0.elapsed_time(a1) / 20 * 1e3, dev)
    out = {"metric": "training_pairs_per_sec", "value": world * B * steps / secs, "unit": UNIT, "n_gpus": world,
           "steps": steps, "warmup": warm, "ms_per_step": secs / steps * 1e3, "mlp_path": mlp,
           "workload": f"C5: training step from keypoints + random descriptors ({NKP} per image, {DESC}-d): mutual-NN "
                       f"matching -> {N} matches + quality, DeepFNet depth 5, {B} pairs/GPU, F-loss + q/t pose loss "
                       "(one fused device head: fepe_b200.losses.deepf_training_loss), MLP weight gradients accumulated in "
                       "place into ONE flat buffer, one all-reduce of it (NCCL), Adam",
           "global_batch": world * B,
           "mean_matches_per_pair": float(torch.stack(n_matches[-steps:]).float().mean()),
           "ms_breakdown": br, "launch": launch, "grad_bytes_allreduced": gbytes,
           "allreduce_alone_us": ar_us,
           "allreduce_busbw_gbs": None if not ar_us else gbytes * 2 * (world - 1) / world / (ar_us * 1e-6) / 1e9}
    del net, opt, flat
    return out


def run_c5(args):
    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__ as entry
    entry.build()
    steps, warm = min(args.steps, 30), max(3, min(args.warmup, 5))
    res = c5_measure(args, dev, rank, world, steps, warm, args.mlp)
    ref = (c5_measure(args, dev, rank, world, max(3, steps // 3), 2, "torch")
           if args.mlp != "torch" and not args.no_extras else None)
    line = {"metric": "training_pairs_per_sec", "value": res["value"], "unit": UNIT, "n_gpus": world,
            "steps": steps, "warmup": warm, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"tc32": "MLP: split-fp16 (3 x tcgen05 kind::f16, fp32 accumulate, fp32 activations) = fp32 parity; "
                              "solver f32 / f64", "bf16": "bf16 MLP forward + backward on tcgen05 (fp32 accumulate) + f32/f64 "
                                                         "solver kernels", "torch": "fp32 library MLP + f32/f64 solver kernels"}[args.mlp],
            "data": "synthetic", "config": {"workload": res["workload"], "global_batch": res["global_batch"]},
            "training_step": res, "fp32_library_mlp": ref}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(line)


def main():
    # stdout must carry exactly ONE JSON line: libraries (NCCL's version banner, torchrun notices) write to file
    # descriptor 1 as well, so fd 1 is pointed at stderr for the whole run and the line goes out through a private copy
    global _OUT
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "C4":
        run_c4(args)
    elif args.workload == "C5":
        run_c5(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
