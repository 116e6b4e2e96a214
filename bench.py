#!/usr/bin/env python
"""bench.py -- measures BASELINE.json's metrics on B200: images/sec of vgg_small Detector:detect on synthetic 800x450
frames (the headline line) and, nested under "also" in the same JSON line, the other configurations BASELINE.json names:
NMS boxes/sec on the 1 M x 21-class sweep, vgg_small lossAndGradient 800x450 x 8 frames, vgg_large detect 1000x600.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA library through the C ABI)
  python bench.py --impl reference --gpus N ...            the reference's CPU path (oracle restatement: the
                                                           reference is pure Lua/Torch7 and cannot run here)
  --workload detect|nms|train                              one configuration alone (default: all, detect as headline)

A step = one pass of the hot path over one batch of synthetic input (detect: one Detector:detect of one frame batch;
nms: one segmented NMS of the sweep; train: one lossAndGradient over the frame batch).  Every timed region is the K
steps bracketed by barrier + synchronize and one CUDA-event pair; the region is REPEATED until at least 0.5 s of device
time has been measured (at least 3 times) and the line reports the median region with min / max (`spread`).
`value` = units of all ranks / median region time with inputs resident in HBM; `e2e` = the same through the public call
with page-locked HOST inputs (H2D of every step's inputs and D2H of its results inside the region).  The detect
configuration keeps `--in-flight` steps in flight (frcnn_detect_begin / frcnn_detect_end on that many library contexts);
`config.sync` is one step at a time.  Rank 0 prints ONE JSON line on stdout; everything else (NCCL INFO included) goes
to stderr."""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: keep the real stdout aside and point fd 1 (where NCCL's INFO lines and any stray
# print go) at stderr
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    # the driver reads the communicator's rank count from NCCL's own log: INFO unless the caller asked for more
    if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
        os.environ["NCCL_DEBUG"] = "INFO"

import numpy as np  # noqa: E402
import torch  # noqa: E402

MIN_REGION_S = 0.5


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="default: 200 (detect), 20 (nms, train)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "detect", "nms", "train"])
    ap.add_argument("--batch", type=int, default=0, help="frames per step per GPU (detect: 1 = configs[1]; train: 8 = configs[2])")
    ap.add_argument("--model", default="vgg_small", choices=["vgg_small", "vgg_large"])
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--nms-n", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--in-flight", type=int, default=12,
                    help="detect: frames (steps) in flight, one library context each (1 = every step synchronous)")
    ap.add_argument("--schedule", default="throughput", choices=["throughput", "latency"],
                    help="detect: launch schedule of the in-flight contexts (frcnn_set_schedule)")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16"],
                    help="evaluate-mode operand format (frcnn_set_eval_precision); training is always bf16")
    ap.add_argument("--min-seconds", type=float, default=MIN_REGION_S, help="device time to accumulate per measured leg")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], bf16_burst=p["bf16_tflops"], bf16_sustained=p["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        self.rows = []
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower() == "active" for r in self.rows)]
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


def model_setup(model, height=0, width=0):
    from oracle import model as OM
    if model == "vgg_small":
        desc, cfg = OM.VGG_SMALL, OM.CFG_DUPLO
        h, w = height or 450, width or 800
    else:
        desc, cfg = OM.VGG_LARGE, OM.CFG_IMAGENET
        h, w = height or 600, width or 1000
    # random weights give ~no detections above 0.95, so the head biases are shifted until
    # ~1-5 % of the anchors pass and the class head is confident: realistic decode / NMS / ROI / cnet work.
    params = OM.detecting_params(OM.init_params(desc, cfg, seed=0, randomize_aux=True))
    return desc, cfg, params, h, w


def conv_flops_per_image(desc, h, w):
    """Algorithmic conv FLOPs of pnet forward: 2*Cin*Cout*k^2*Hout*Wout per conv."""
    total, cin, dims = 0.0, 3, []
    for l in desc["layers"]:
        for _ in range(l["conv_steps"]):
            total += 2.0 * cin * l["filters"] * l["kW"] * l["kH"] * h * w
            cin = l["filters"]
        h, w = (h + 1) // 2, (w + 1) // 2
        dims.append((h, w, cin))
    for a in desc["anchor_nets"]:
        ih, iw, c = dims[a["input"] - 1]
        oh, ow = ih - a["kW"] + 1, iw - a["kW"] + 1
        total += 2.0 * c * a["n"] * a["kW"] ** 2 * oh * ow + 2.0 * a["n"] * 18 * oh * ow
    return total


def detect_workload(model, w, h, batch):
    cfgname = {("vgg_small", 1): " (BASELINE configs[1])", ("vgg_large", 1): " (BASELINE configs[3], per GPU)"}.get((model, batch), "")
    return "%s Detector:detect %dx%d batch=%d per GPU%s" % (model, w, h, batch, cfgname)


def nms_workload(n):
    return "nms.lua sweep: %d boxes in 21 class segments, overlap 0.25, key y2, sort included (BASELINE configs[4])" % n


def train_workload(model, w, h, batch):
    return ("%s lossAndGradient (objective.lua) %dx%d, %d frames per GPU, 128+128 anchors, 8 boxes per frame "
            "(BASELINE configs[2])" % (model, w, h, batch))


# ----------------------------------------------------------------------------------------------- CPU baselines (oracle)
def cpu_detect(desc, cfg, params, h, w, budget_s=10.0, max_frames=40):
    """The restated reference CPU path (PyTorch-CPU fp32 = TH/THNN descendants + the Lua loops in Python/numpy), all host
    threads, on a bounded sample.  Returns (cpu_baseline dict, fp32 result of the first frame for the precision report)."""
    from oracle import detector as OD, model as OM
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    od = OD.Detector(desc, cfg, params)
    f = OM.synthetic_frame(h, w, seed=0)
    first = od.detect(f, return_intermediates=True)
    n, t0 = 0, time.perf_counter()
    while n < 3 or (time.perf_counter() - t0 < budget_s and n < max_frames):
        od.detect(f)
        n += 1
    dt = time.perf_counter() - t0
    return dict(value=n / dt, unit="images/s", cores=cores, kind="port",
                sample="%d frames %dx%d through the oracle (PyTorch-CPU fp32 + restated Lua loops), %d threads" % (n, w, h, cores)), (f, first)


def cpu_nms(n_total, budget_s=5.0):
    from oracle import boxes as OB, nms_c
    cores = os.cpu_count() or 1
    n = min(n_total, 200_000)
    b = OB.sweep_boxes(n, seed=0)
    perm, seg = OB.class_segments(n, 21, seed=0)
    b = b[perm]
    nms_c.pair_iou_count()
    t0 = time.perf_counter()
    reps = 0
    while reps < 1 or time.perf_counter() - t0 < budget_s:
        nms_c.nms_segmented(b, seg, 0.25, threads=cores)
        reps += 1
    dt = time.perf_counter() - t0
    pairs = nms_c.pair_iou_count()
    return dict(value=n * reps / dt, unit="boxes/s", cores=min(cores, 21), kind="port", pair_iou_per_sec=pairs / dt,
                sample="%d boxes in 21 class segments x %d reps, C restatement of nms.lua, one thread per class" % (n, reps))


def cpu_train(desc, cfg, params, h, w, budget_s=15.0):
    """lossAndGradient of the oracle (PyTorch-CPU fp32 autograd + restated Lua loops), frame by frame as the reference."""
    from oracle import anchors as OA, model as OM, objective as OO
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    oa = OA.Anchors(desc["layers"], desc["anchor_nets"], cfg["scales"])
    img = OM.synthetic_frame(h, w, seed=0)
    with torch.no_grad():
        dims = [tuple(o.shape) for o in OM.pnet_forward(desc, params, img)]
    pos, neg, _ = OO.synthetic_examples(oa, dims, w, h, 128, 128, 8, cfg["class_count"], seed=0)
    g = torch.Generator().manual_seed(0)
    dm = {"b%d_c1" % (i + 1): (torch.rand(l["filters"], generator=g) > l["dropout"]).float()
          for i, l in enumerate(desc["layers"]) if l["dropout"]}
    R = len(OO.clean_anchors(pos, dims)) + len(OO.clean_anchors(neg, dims))
    cm = {"fc%d" % (i + 1): (torch.rand(R, l["n"], generator=g) > l["dropout"]).float() for i, l in enumerate(desc["class_layers"])}
    n, t0 = 0, time.perf_counter()
    while n < 2 or (time.perf_counter() - t0 < budget_s and n < 20):
        OO.loss_and_gradient_image(desc, cfg, params, img, pos, neg, dropout_masks=dm, cnet_masks=cm)
        n += 1
    dt = time.perf_counter() - t0
    return dict(value=n / dt, unit="images/s", cores=cores, kind="port",
                sample="%d frames %dx%d (128+128 anchors) through the oracle objective (PyTorch-CPU fp32 autograd), %d threads" % (n, w, h, cores))


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own CPU implementation of the path.  It is pure Lua on Torch7 (no Lua toolchain in this
    image, and its path needs cunn), so the timed code is the oracle restatement: PyTorch-CPU fp32 for pnet / cnet
    (the TH/THNN descendants) + the Lua-loop logic restated in Python/numpy, all host threads."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = "detect" if args.workload == "all" else args.workload
    warm = max(0, min(args.warmup, 1))
    if wl == "nms":
        from oracle import boxes as OB, nms_c
        n = min(args.nms_n, 200_000)
        b = OB.sweep_boxes(n, seed=0)
        perm, seg = OB.class_segments(n, 21, seed=0)
        b = b[perm]
        for _ in range(warm):
            nms_c.nms_segmented(b, seg, 0.25, threads=cores)
        steps = max(1, min(args.steps or 5, 5))
        t0 = time.perf_counter()
        for _ in range(steps):
            nms_c.nms_segmented(b, seg, 0.25, threads=cores)
        dt = time.perf_counter() - t0
        v, unit, metric = n * steps / dt, "boxes/s", "nms_boxes_per_sec"
        workload = nms_workload(args.nms_n)
        sample = "%d boxes x 21 classes per step (bounded sample of the sweep), C restatement of nms.lua, one thread per class" % n
    elif wl == "train":
        desc, cfg, params, h, w = model_setup(args.model, args.height, args.width)
        steps = max(1, min(args.steps or 3, 3))
        t0 = time.perf_counter()
        cb = cpu_train(desc, cfg, params, h, w, budget_s=10.0 * steps)
        dt = time.perf_counter() - t0
        v, unit, metric = cb["value"], "images/s", "train_images_per_sec"
        workload = train_workload(args.model, w, h, args.batch or 8)
        sample = cb["sample"]
    else:
        from oracle import detector as OD, model as OM
        desc, cfg, params, h, w = model_setup(args.model, args.height, args.width)
        od = OD.Detector(desc, cfg, params)
        frames = [OM.synthetic_frame(h, w, seed=s) for s in range(2)]
        # bounded sample: at most 40 frames (about 6 s on 16 cores) whatever --steps says
        steps = max(1, min(args.steps or 20, 40))
        for i in range(warm):
            od.detect(frames[i % 2])
        t0 = time.perf_counter()
        for i in range(steps):
            od.detect(frames[i % 2])
        dt = time.perf_counter() - t0
        v, unit, metric = steps / dt, "images/s", "images_per_sec"
        workload = detect_workload(args.model, w, h, args.batch or 1)
        sample = "%d frames %dx%d, PyTorch-CPU fp32 oracle, %d threads" % (steps, w, h, cores)
    line = dict(metric=metric, value=v, unit=unit, n_gpus=args.gpus, steps=steps, warmup=warm, ms_per_step=1e3 * dt / steps,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=workload,
                            note="CPU restatement of the Lua/Torch7 path (the reference itself cannot execute here: no Lua/Torch7, "
                                 "its path requires cunn); each step is a bounded sample of the workload"),
                cpu_baseline=dict(value=v, unit=unit, cores=cores, kind="port", sample=sample),
                e2e=dict(value=v, unit=unit, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    emit(line)


# ----------------------------------------------------------------------------------------------- our arm
class Harness:
    """Device, distributed plumbing and the repeated timed region shared by the workloads."""

    def __init__(self, args):
        self.args = args
        self.rank, self.world, self.local = dist_env()
        assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm"
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        import frcnn_b200 as F
        self.F = F
        self.pk = peaks()
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
        self.sampler = ClockSampler(self.local)

    def barrier(self):
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x):
        return self.F.reduce_timing(x, 0, self.world, self.dist, "cuda")[0]

    def sum_over_ranks(self, x):
        return self.F.reduce_timing(0, x, self.world, self.dist, "cuda")[1]

    def region(self, run_k_steps, warm):
        """One timed region: warm-up, barrier + synchronize, ONE event pair around exactly K steps, barrier + synchronize;
        max over ranks.  Repeated until MIN_REGION_S of device time has been accumulated (>= 3 regions).  Returns
        (median ms, dict(min, median, max, repeats, total_s))."""
        warm()
        times = []
        total, target = 0.0, self.args.min_seconds * 1e3
        while len(times) < 3 or total < target:
            self.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()   # legacy default stream: ordered against the (blocking) streams of the library contexts
            run_k_steps()
            e1.record()
            self.barrier()
            ms = self.max_over_ranks(e0.elapsed_time(e1))
            times.append(ms)
            total += ms
            if len(times) >= 400:
                break
        med = float(np.median(times))
        return med, dict(min_ms=float(min(times)), median_ms=med, max_ms=float(max(times)), repeats=len(times), total_s=total * 1e-3)

    def flushed_steps(self, fn, steps, warmup):
        """K synchronous steps with the L2 flushed before each (outside its event pair): sum of the K event pairs."""
        for _ in range(warmup):
            fn()
        self.barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in evs:
            self.flush.fill_(1)
            a.record()
            fn()
            b.record()
        self.barrier()
        return self.max_over_ranks(sum(a.elapsed_time(b) for a, b in evs))

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def bench_detect(hs, model, steps, warmup, batch, with_cpu, headline):
    args, F = hs.args, hs.F
    from oracle import model as OM  # weight / frame generators (bench infrastructure) + the cpu_baseline leg
    ffi, L = F.ffi, F.lib()
    desc, cfg, params, h, w = model_setup(model, args.height if headline else 0, args.width if headline else 0)
    fac, fcfg = (F.vgg_small, F.duplo_cfg) if model == "vgg_small" else (F.vgg_large, F.imgnet_cfg)
    m = fac(fcfg, device=hs.local)
    m.load_params(params)
    m.set_eval_precision(args.precision)
    det = F.Detector(m)
    B = batch
    frames = torch.stack([OM.synthetic_frame(h, w, seed=100 * hs.rank + s) for s in range(B)])
    frames_dev = frames.cuda()
    n_det = ffi.new("int*")
    out = det._out

    def step_dev():
        rc = L.frcnn_detect_dev(m.ctx, ffi.cast("const float*", frames_dev.data_ptr()), B, h, w, out, det._cap, n_det)
        if rc != 0:
            raise RuntimeError(ffi.string(L.frcnn_last_error(m.ctx)).decode())

    step_dev()
    stats = det.stats()
    # ---- synchronous steps (one frame batch, wait for its winners, L2 flushed in between): the latency figure
    sync_steps = min(steps, 20)
    ms_sync = hs.flushed_steps(step_dev, sync_steps, warmup)
    # ---- the measured configuration: `in_flight` steps in flight (frcnn_detect_begin / frcnn_detect_end on S library
    # contexts), every step a different resident frame batch out of a set larger than L2
    S = max(1, args.in_flight)
    pipe = F.DetectorPipeline(m, in_flight=S, schedule=args.schedule)
    ctxs = [mm.ctx for mm in pipe.models]
    n_sets = max(S + 1, -(-(140 << 20) // (B * 3 * h * w * 4)))  # > 126 MB of L2 in total
    sets_dev = [frames_dev.roll(shifts=(3 * i, 7 * i), dims=(-2, -1)).contiguous() for i in range(n_sets)]
    sets_host = [t.cpu().pin_memory() for t in sets_dev]             # e2e leg: page-locked host frames
    ptr_dev = [ffi.cast("const float*", t.data_ptr()) for t in sets_dev]
    ptr_host = [ffi.cast("const float*", t.data_ptr()) for t in sets_host]
    cursor = dict(i=0)

    def run_steps(k, ptrs, on_dev):
        busy = [False] * S
        base = cursor["i"]
        for i in range(k):
            sl = i % S
            if busy[sl]:
                rc = L.frcnn_detect_end(ctxs[sl], out, det._cap, n_det)
                assert rc == 0, ffi.string(L.frcnn_last_error(ctxs[sl])).decode()
            rc = L.frcnn_detect_begin(ctxs[sl], ptrs[(base + i) % n_sets], on_dev, B, h, w)
            assert rc == 0, ffi.string(L.frcnn_last_error(ctxs[sl])).decode()
            busy[sl] = True
        for j in range(S):
            sl = (k + j) % S
            if busy[sl]:
                rc = L.frcnn_detect_end(ctxs[sl], out, det._cap, n_det)
                assert rc == 0, ffi.string(L.frcnn_last_error(ctxs[sl])).decode()
        cursor["i"] = base + k

    l0 = sum(mm.launch_count() for mm in pipe.models)
    s0 = dict(n=0)

    def warm_dev():
        run_steps(max(warmup, 3) * S, ptr_dev, 1)   # every context has captured and replayed its graph
        s0["n"] += max(warmup, 3) * S

    def k_dev():
        run_steps(steps, ptr_dev, 1)
        s0["n"] += steps

    hs.sampler.start()
    ms, spread = hs.region(k_dev, warm_dev)
    clocks = hs.sampler.stop()
    launches = sum(mm.launch_count() for mm in pipe.models) - l0
    launches_per_step = launches // max(s0["n"], 1)
    ms_e2e, spread_e2e = hs.region(lambda: run_steps(steps, ptr_host, 0), lambda: run_steps(max(warmup, 3) * S, ptr_host, 0))
    # counters (16 + 2 per frame ints) + the winners (the first 256 are copied speculatively with the counters)
    d2h = 4 * (16 + 2 * B) + max(256, int(n_det[0])) * ffi.sizeof("frcnn_detection")
    # roofline pass: synchronous steps with an event pair around every conv/GEMM launch
    L.frcnn_set_profiling(m.ctx, 1)
    step_dev()
    conv_ms = conv_fl = 0.0
    conv_n = 0
    stage_ms = np.zeros(6)
    pm, pf, pn, st = ffi.new("float*"), ffi.new("double*"), ffi.new("int*"), ffi.new("float[6]")
    prof_steps = min(steps, 20)
    for _ in range(prof_steps):
        hs.flush.fill_(1)
        step_dev()
        L.frcnn_last_conv_profile(m.ctx, pm, pf, pn)
        L.frcnn_last_timings(m.ctx, st)
        conv_ms += pm[0]
        conv_fl += pf[0]
        conv_n += pn[0]
        stage_ms += np.array([st[i] for i in range(6)])
    L.frcnn_set_profiling(m.ctx, 0)
    achieved_sync = conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    flops_per_step = conv_fl / max(prof_steps, 1)
    # pipelined region: launches of different frames overlap, so per-launch event pairs no longer isolate the kernel;
    # the conservative figure is the algorithmic conv/GEMM FLOPs of the K steps over the WHOLE timed region (every other
    # kernel of the step included in the denominator): a lower bound of the conv kernels' own rate
    achieved = flops_per_step * steps / (ms * 1e-3) / 1e12
    traffic, traffic_src = None, None
    try:
        summ = json.load(open(os.path.join(ROOT, "profiles", "r2_summary.json")))
        rows = summ.get("full_%s_b%d" % (model, B))
        if rows and (h, w) == ((450, 800) if model == "vgg_small" else (600, 1000)):
            traffic = sum(r["dram_mb"] for r in rows) * 1e6
            traffic_src = "profiles/r2_ncu_full_b%d.md (dram__bytes_read.sum + dram__bytes_write.sum over the %d tcgen05 launches of one %s step)" % (B, len(rows), model)
    except Exception:
        pass
    total_frames = hs.sum_over_ranks(float(B)) * steps
    peak = hs.pk["bf16_sustained"]
    line = dict(metric="images_per_sec", value=total_frames / (ms * 1e-3), unit="images/s", n_gpus=hs.world, steps=steps, warmup=warmup,
                ms_per_step=ms / steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f16" if args.precision == "fp16" else "bf16", data="synthetic",
                config=dict(workload=detect_workload(model, w, h, B),
                            weights="seeded random init, head biases shifted so the detector stages have work (SURVEY 8d)",
                            operands="%s tensor-core operands, fp32 accumulation (reference: fp32); agreement with the fp32 path: `precision`" % args.precision,
                            stages=stats, in_flight=S,
                            l2="every step reads a different resident frame batch out of %d (%.0f MB > L2); no flush inside "
                               "the timed region" % (n_sets, n_sets * B * 3 * h * w * 4 / 2 ** 20),
                            sync=dict(ms_per_step=ms_sync / sync_steps, images_per_sec=B * sync_steps / (ms_sync * 1e-3),
                                      note="one step at a time (frcnn_detect_dev), L2 flushed between steps"),
                            sharding="frames over ranks, no collective", launches_per_step=int(launches_per_step)),
                spread=spread, clocks=clocks,
                e2e=dict(value=total_frames / (ms_e2e * 1e-3), unit="images/s", h2d_bytes_per_step=int(frames.numel() * 4),
                         d2h_bytes_per_step=int(d2h), in_flight=S, spread=spread_e2e,
                         note="frcnn_detect_begin on page-locked host frames (H2D inside), frcnn_detect_end reads the winners"),
                gpu_launches=int(launches_per_step * steps),
                roofline=dict(bound="tensor", kernel="conv_halo_kernel / conv_igemm_kernel / conv_first_kernel (all tcgen05 launches of a step)",
                              achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak if peak else None,
                              traffic=traffic, traffic_source=traffic_src,
                              how="algorithmic conv/GEMM FLOPs of the K steps / device time of the whole timed region (%d steps in "
                                  "flight: launches overlap, so the whole step is the denominator -- a lower bound of the kernels' "
                                  "own rate)" % S,
                              flops_per_step=flops_per_step,
                              pnet_conv_flops_per_image=conv_flops_per_image(desc, h, w),
                              sync_pass=dict(achieved=achieved_sync, frac=achieved_sync / peak if peak else None,
                                             launches_per_step=conv_n // max(prof_steps, 1), ms_per_step=conv_ms / max(prof_steps, 1),
                                             stage_ms_per_step=dict(zip(["pnet", "decode+nms", "roi_pool", "cnet", "final_nms", "total"],
                                                                        (stage_ms / max(prof_steps, 1)).round(4).tolist())),
                                             how="one step at a time, an event pair around every tcgen05 launch"),
                              peak_source=hs.pk["source"] + " bf16 sustained (cuBLAS; fp16 operands run at the same rate)"))
    if hs.rank == 0 and with_cpu and hs.world == 1:
        from oracle import compare as OC
        cb, (f0, fp32) = cpu_detect(desc, cfg, params, h, w, budget_s=10.0 if headline else 6.0)
        line["cpu_baseline"] = cb
        # precision contract, measured in this run: the CUDA path against the pure fp32 oracle from the same frame
        winners = det.detect(f0.numpy())
        maps = [o.cpu() for o in m.pnet.forward(f0.cuda())]
        rep = OC.precision_report(desc, cfg, params, f0, maps, winners, fp32_result=fp32,
                                  quant=OM.fp16_round if args.precision == "fp16" else OM.bf16_round)
        line["precision"] = dict(vs="pure fp32 oracle (restated reference) from the same frame", operands=args.precision,
                                 matches=rep["matches"], candidates=rep["candidates"], winners=rep["winners"],
                                 map_max_abs_err={r["map"]: r["max_abs"] for r in rep["maps"]},
                                 fg_margin_max_abs_err=max(r["margin_max_abs"] for r in rep["maps"][:4]))
    pipe.close()
    m.close()
    return line


def bench_nms(hs, steps, warmup, with_cpu):
    args, F = hs.args, hs.F
    from oracle import boxes as OB
    n = args.nms_n
    # shard the 21 class segments over the ranks: no collective
    perm, seg = OB.class_segments(n, 21, seed=0)
    boxes = OB.sweep_boxes(n, seed=0)[perm]
    mine = F.shard_segments(21, hs.world, hs.rank)
    parts = [boxes[seg[s]:seg[s + 1]] for s in mine]
    local_boxes = np.ascontiguousarray(np.concatenate(parts)) if parts else np.zeros((0, 4), np.float32)
    pinned_boxes = torch.from_numpy(local_boxes).pin_memory()
    local_boxes = pinned_boxes.numpy()
    local_seg = np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.int64)
    m = F.vgg_small(F.duplo_cfg, device=hs.local)
    bd = torch.from_numpy(local_boxes).cuda()
    l0 = m.launch_count()
    picked = dict(n=0)
    calls = dict(n=0)

    def step_dev():
        F.nms_segmented_dev(bd, local_seg, 0.25, model=m)
        calls["n"] += 1

    def step_host():
        _, cnts = F.nms_segmented(local_boxes, local_seg, 0.25, model=m)
        picked["n"] = int(cnts.sum())

    def k_of(fn):
        def run():
            for _ in range(steps):
                hs.flush.fill_(1)   # L2 flush between the steps (the boxes are 16 MB: they would stay resident)
                fn()
        return run

    hs.sampler.start()
    ms, spread = hs.region(k_of(step_dev), lambda: [step_dev() for _ in range(max(warmup, 3))])
    clocks = hs.sampler.stop()
    launches = (m.launch_count() - l0) // max(calls["n"], 1)
    ms_e2e, spread_e2e = hs.region(k_of(step_host), lambda: [step_host() for _ in range(max(warmup // 2, 3))])
    # the flush itself (a 256 MB fill, ~0.04 ms) sits inside the region: measured separately and subtracted
    ms_flush, _ = hs.region(lambda: [hs.flush.fill_(1) for _ in range(steps)], lambda: None)
    ms, ms_e2e = max(ms - ms_flush, 1e-6), max(ms_e2e - ms_flush, 1e-6)
    total = hs.sum_over_ranks(float(len(local_boxes)))
    value = total * steps / (ms * 1e-3)
    line = dict(metric="nms_boxes_per_sec", value=value, unit="boxes/s", n_gpus=hs.world, steps=steps, warmup=warmup,
                ms_per_step=ms / steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=nms_workload(n), l2="flushed between steps (the flush's own time subtracted)",
                            sharding="class segments round-robin over ranks, no collective", launches_per_step=int(launches)),
                spread=spread, clocks=clocks,
                e2e=dict(value=total * steps / (ms_e2e * 1e-3), unit="boxes/s", h2d_bytes_per_step=int(local_boxes.nbytes),
                         d2h_bytes_per_step=int(8 * picked["n"] + 8 * len(mine)), spread=spread_e2e,
                         note="boxes from page-locked host memory; the counts and the picked indices of every class come back"),
                gpu_launches=int(launches * steps))
    if hs.rank == 0:
        # pair-IoU evaluations of the reference algorithm on this input (every pick tests all remaining boxes,
        # nms.lua:72-96): counted by the C restatement on a 200 k-box sample and per-box-scaled -- NMS is dependency /
        # ALU-bound, the HBM roofline (16 B/box + 4 B class id) is reported for completeness only
        cb = cpu_nms(n) if (with_cpu and hs.world == 1) else None
        hbm = 20.0 * value / 1e9 / hs.world
        line["roofline"] = dict(bound="hbm", kernel="nms_cta_kernel / nms_mask_kernel / radix sort passes", achieved=hbm, peak=hs.pk["hbm"],
                                unit="GB/s", frac=hbm / hs.pk["hbm"], traffic=None, peak_source=hs.pk["source"],
                                note="algorithmic bytes only (16 B/box read + 4 B class id); greedy NMS is bound by the dependency chain "
                                     "of its picks and by ALU work, see pair_iou_per_sec")
        if cb:
            line["cpu_baseline"] = cb
            pairs_per_box = cb["pair_iou_per_sec"] / cb["value"]
            line["pair_iou_per_sec"] = dict(value=value * pairs_per_box, reference_pairs_per_box=pairs_per_box,
                                            note="reference-equivalent pair-IoU evaluations (what nms.lua would compute for these boxes) per "
                                                 "second of GPU time; the GPU computes a superset (1024 x 1024 bit-matrix rounds)")
    m.close()
    return line


def bench_train(hs, model, steps, warmup, batch, with_cpu):
    # BASELINE configs[2]: vgg_small training fwd/bwd (objective.lua) 800x450, batch frames per GPU, 128 positive +
    # 128 negative anchors and 8 ground-truth boxes per frame; frames sharded over the ranks,
    # ONE all-reduce of the flat gradient (+ counters) per step
    args, F = hs.args, hs.F
    from oracle import anchors as OA, model as OM, objective as OO
    desc, cfg, params, h, w = model_setup(model)
    m = (F.vgg_small if model == "vgg_small" else F.vgg_large)(F.duplo_cfg if model == "vgg_small" else F.imgnet_cfg, device=hs.local)
    m.load_params(params)
    oa = OA.Anchors(desc["layers"], desc["anchor_nets"], cfg["scales"])
    dims = m.output_dims(h, w)
    B = max(batch, 1)
    batch_list, host_frames = [], []
    for s in range(B):
        pos, neg, _ = OO.synthetic_examples(oa, dims, w, h, 128, 128, 8, cfg["class_count"], seed=1000 * hs.rank + s)
        pos, neg = F.clean_anchors(pos, dims), F.clean_anchors(neg, dims)
        f = OM.synthetic_frame(h, w, seed=100 * hs.rank + s)
        host_frames.append(f.pin_memory())
        # the example records are marshalled once by the batch iterator (BatchIterator:nextTraining's job in the
        # reference), not inside the timed lossAndGradient call
        batch_list.append(dict(img=f.cuda(), positive=pos, negative=neg, packed=(m.pack_examples(pos), m.pack_examples(neg))))
    objective = F.create_objective(m, hs.dist)
    state = dict(step=0, loss=0.0)
    l0 = m.launch_count()

    def step():
        state["step"] += 1
        loss, _, _ = objective(batch_list, seed=state["step"])
        state["loss"] = loss

    prefetch = F.FramePrefetcher(device=hs.local, depth=2)
    host_stack = torch.stack(host_frames).pin_memory()    # the step's frames, page-locked, one asynchronous copy

    def upload_next():
        prefetch.submit(host_stack)

    def step_e2e():
        # objective.lua:66: img = batch[i].img:cuda() -- every frame of the step comes from page-locked host memory (one
        # asynchronous copy per frame on the prefetcher's stream, enqueued one step ahead: BatchIterator's job); the loss
        # scalars come back (the objective reads them on the host)
        if prefetch.pending() == 0:
            upload_next()
        stack = prefetch.get()
        for i, b in enumerate(batch_list):
            b["img"] = stack[i]
        upload_next()                      # the NEXT step's frames go up while this step computes
        step()
        if state.get("stack") is not None:
            prefetch.recycle(state["stack"])
        state["stack"] = stack

    def k_of(fn):
        def run():
            for _ in range(steps):
                fn()
        return run

    hs.sampler.start()
    ms, spread = hs.region(k_of(step), lambda: [step() for _ in range(max(warmup, 3))])
    clocks = hs.sampler.stop()
    launches = (m.launch_count() - l0) // max(state["step"], 1)
    ms_e2e, spread_e2e = hs.region(k_of(step_e2e), lambda: [step_e2e() for _ in range(3)])
    collective = None
    if hs.world > 1:
        # what the gradient all-reduce costs a step: the same steps with the collective left out (every rank on its own)
        local_objective = F.create_objective(m, None)

        def step_local():
            state["step"] += 1
            local_objective(batch_list, seed=state["step"])

        ms_local, _ = hs.region(k_of(step_local), lambda: [step_local() for _ in range(3)])
        info = m.dp_info() if hasattr(m, "dp_info") else {}
        collective = dict(kind="ncclAllReduce(sum, fp32) in place on the flat gradient + counters, bucketed cnet -> heads -> block 4..1, "
                               "launched from inside pnet:backward as each bucket's last wgrad finishes (frcnn_dp_*)",
                          bytes_per_step=int(m.gradient.numel() * 4 + 32), ms_per_step_with=ms / steps, ms_per_step_without=ms_local / steps,
                          exposed_ms_per_step=(ms - ms_local) / steps, **info)
    total = hs.sum_over_ranks(float(B)) * steps
    fwd = conv_flops_per_image(desc, h, w)
    grad_bytes = int(m.gradient.numel() * 4)
    line = dict(metric="train_images_per_sec", value=total / (ms * 1e-3), unit="images/s", n_gpus=hs.world, steps=steps, warmup=warmup,
                ms_per_step=ms / steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
                config=dict(workload=train_workload(model, w, h, B),
                            l2="every step streams %d frames + %.0f MB of activations (> L2); no flush" % (B, 1100.0 * B / 8),
                            sharding="frames over ranks; the flat gradient (%.1f MB + 7 counters) is all-reduced every step" % (grad_bytes / 1e6),
                            launches_per_step=int(launches), last_loss=float(state["loss"])),
                spread=spread, clocks=clocks, gpu_launches=int(launches * steps),
                e2e=dict(value=total / (ms_e2e * 1e-3), unit="images/s", h2d_bytes_per_step=int(B * 3 * h * w * 4 + B * 256 * 96),
                         d2h_bytes_per_step=int(B * 16 + 28), spread=spread_e2e,
                         note="lossAndGradient with every frame copied from page-locked host memory inside the timed region (objective.lua:66; "
                              "FramePrefetcher: asynchronous copies on a side stream, one step ahead) plus the example records; the loss "
                              "sums and counters come back"),
                roofline=dict(bound="tensor", kernel="conv_igemm_kernel / conv_halo_kernel / conv_wgrad_halo_kernel (fwd + dgrad + wgrad launches of a step)",
                              achieved=3 * fwd * B / (ms / steps * 1e-3) / 1e12, peak=hs.pk["bf16_sustained"], unit="TFLOP/s",
                              frac=3 * fwd * B / (ms / steps * 1e-3) / 1e12 / hs.pk["bf16_sustained"], traffic=None,
                              peak_source=hs.pk["source"] + " bf16 sustained",
                              note="3 x forward conv FLOPs per frame over the whole step (every other kernel and the all-reduce in the "
                                   "denominator): lower bound of the tcgen05 kernels' own rate"))
    prefetch.close()
    if collective:
        line["collective"] = collective
    if hs.rank == 0 and with_cpu and hs.world == 1:
        line["cpu_baseline"] = cpu_train(desc, cfg, params, h, w)
    m.close()
    return line


def run_b200(args):
    hs = Harness(args)
    cpu = not args.no_cpu_baseline
    wl = args.workload
    if wl == "nms":
        line = bench_nms(hs, args.steps or 20, args.warmup, cpu)
    elif wl == "train":
        line = bench_train(hs, args.model, args.steps or 20, args.warmup, args.batch or 8, cpu)
    else:
        line = bench_detect(hs, args.model, args.steps or 200, args.warmup, args.batch or 1, cpu, headline=True)
        if wl == "all":
            # the other configurations BASELINE.json names, each measured like the headline and nested in the same line
            also = {}
            also["nms"] = bench_nms(hs, 20, args.warmup, cpu)
            also["train"] = bench_train(hs, "vgg_small", 20, args.warmup, 8, cpu)
            if args.model == "vgg_small":
                also["vgg_large"] = bench_detect(hs, "vgg_large", 100, args.warmup, 1, cpu, headline=False)
            line["also"] = also
    if hs.rank == 0:
        emit(line)
    hs.close()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
