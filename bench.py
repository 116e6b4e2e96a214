#!/usr/bin/env python
"""bench.py -- measures BASELINE.json's metric: images/sec of vgg_small Detector:detect on synthetic 800x450 frames
(and, with --workload nms, NMS boxes/sec on the 21-class sweep).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA library through the C ABI)
  python bench.py --impl reference --gpus N ...            the reference's CPU path (oracle restatement: the
                                                           reference is pure Lua/Torch7 and cannot run here)

A step = one Detector:detect pass over one batch of synthetic frames (batch 1 = BASELINE configs[1]).  The measured
configuration keeps `--in-flight` steps in flight (frcnn_detect_begin / frcnn_detect_end on that many library contexts,
throughput schedule): `value` = frames of all ranks / device time of the K steps (one CUDA-event pair around the whole
region, every step a different resident frame batch out of a set larger than L2, max over ranks); `e2e` = the same loop
fed from page-locked HOST frames (the H2D copy of every step's frames and the D2H read of its winners inside the timed
region).  `config.sync` is one step at a time through frcnn_detect_dev with the L2 flushed in between (latency).  The
roofline entry is the algorithmic conv/GEMM FLOPs of the K steps over the same timed region (a lower bound of the tcgen05
kernels' own rate); `roofline.sync_pass` keeps the per-launch event-pair numbers of synchronous steps.
One JSON line on stdout (rank 0)."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line

import numpy as np  # noqa: E402
import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="default: 200 (detect), 20 (nms, train)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="detect", choices=["detect", "nms", "train"])
    ap.add_argument("--batch", type=int, default=1, help="frames per step per GPU (configs[1]: batch=1)")
    ap.add_argument("--model", default="vgg_small", choices=["vgg_small", "vgg_large"])
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--nms-n", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--in-flight", type=int, default=6,
                    help="detect: frames (steps) in flight, one library context each (1 = every step synchronous)")
    ap.add_argument("--schedule", default="throughput", choices=["throughput", "latency"],
                    help="detect: launch schedule of the in-flight contexts (frcnn_set_schedule)")
    a = ap.parse_args()
    if a.steps <= 0:
        a.steps = 200 if (a.workload == "detect" and a.impl == "b200") else 20
    return a


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


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


def model_setup(args):
    from oracle import model as OM
    if args.model == "vgg_small":
        desc, cfg = OM.VGG_SMALL, OM.CFG_DUPLO
        h, w = args.height or 450, args.width or 800
    else:
        desc, cfg = OM.VGG_LARGE, OM.CFG_IMAGENET
        h, w = args.height or 600, args.width or 1000
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
    if args.workload == "nms":
        from oracle import boxes as OB, nms_c
        n = min(args.nms_n, 200_000)
        b = OB.sweep_boxes(n, seed=0)
        perm, seg = OB.class_segments(n, 21, seed=0)
        b = b[perm]
        for _ in range(min(args.warmup, 1)):
            nms_c.nms_segmented(b, seg, 0.25, threads=cores)
        t0 = time.perf_counter()
        steps = max(1, min(args.steps, 5))
        for _ in range(steps):
            nms_c.nms_segmented(b, seg, 0.25, threads=cores)
        dt = time.perf_counter() - t0
        v = n * steps / dt
        line = dict(metric="nms_boxes_per_sec", value=v, unit="boxes/s", n_gpus=args.gpus, steps=steps, warmup=min(args.warmup, 1),
                    ms_per_step=1e3 * dt / steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                    data="synthetic", impl="reference",
                    config=dict(workload="nms.lua sweep: %d boxes in 21 class segments, overlap 0.25, key y2" % n),
                    cpu_baseline=dict(value=v, unit="boxes/s", cores=cores, kind="port",
                                      sample="%d boxes x 21 classes, C restatement of nms.lua, one thread per class" % n),
                    e2e=dict(value=v, unit="boxes/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
        print(json.dumps(line))
        return
    from oracle import detector as OD, model as OM
    desc, cfg, params, h, w = model_setup(args)
    od = OD.Detector(desc, cfg, params)
    frames = [OM.synthetic_frame(h, w, seed=s) for s in range(2)]
    steps, warm = max(1, min(args.steps, 8)), min(args.warmup, 1)
    for i in range(warm):
        od.detect(frames[i % 2])
    t0 = time.perf_counter()
    for i in range(steps):
        od.detect(frames[i % 2])
    dt = time.perf_counter() - t0
    v = steps / dt
    line = dict(metric="images_per_sec", value=v, unit="images/s", n_gpus=args.gpus, steps=steps, warmup=warm,
                ms_per_step=1e3 * dt / steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=dict(workload="%s Detector:detect %dx%d batch=1 (CPU restatement of the Lua/Torch7 path)" % (args.model, w, h),
                            note="the reference itself cannot execute here: no Lua/Torch7, path requires cunn"),
                cpu_baseline=dict(value=v, unit="images/s", cores=cores, kind="port",
                                  sample="%d frames %dx%d, PyTorch-CPU fp32 oracle, %d threads" % (steps, w, h, cores)),
                e2e=dict(value=v, unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- our arm
def cpu_baseline_detect(args, desc, cfg, params, h, w):
    from oracle import detector as OD, model as OM
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    od = OD.Detector(desc, cfg, params)
    f = OM.synthetic_frame(h, w, seed=0)
    od.detect(f)
    n, t0 = 0, time.perf_counter()
    while n < 3 or (time.perf_counter() - t0 < 10.0 and n < 40):
        od.detect(f)
        n += 1
    dt = time.perf_counter() - t0
    return dict(value=n / dt, unit="images/s", cores=cores, kind="port",
                sample="%d frames %dx%d through the oracle (PyTorch-CPU fp32 + restated Lua loops), %d threads" % (n, w, h, cores))


def cpu_baseline_nms(n_total):
    from oracle import boxes as OB, nms_c
    cores = os.cpu_count() or 1
    n = min(n_total, 200_000)
    b = OB.sweep_boxes(n, seed=0)
    perm, seg = OB.class_segments(n, 21, seed=0)
    b = b[perm]
    t0 = time.perf_counter()
    reps = 0
    while reps < 1 or time.perf_counter() - t0 < 5.0:
        nms_c.nms_segmented(b, seg, 0.25, threads=cores)
        reps += 1
    dt = time.perf_counter() - t0
    return dict(value=n * reps / dt, unit="boxes/s", cores=min(cores, 21), kind="port",
                sample="%d boxes in 21 class segments x %d reps, C restatement of nms.lua, one thread per class" % (n, reps))


def run_b200(args):
    rank, world, local = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm"
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import frcnn_b200 as F
    from oracle import model as OM  # weight / frame generators only (bench infrastructure)
    ffi, L = F.ffi, F.lib()
    pk = peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        return F.reduce_timing(x, 0, world, dist if world > 1 else None, "cuda")[0]

    def sum_over_ranks(x):
        return F.reduce_timing(0, x, world, dist if world > 1 else None, "cuda")[1]

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in evs:
            flush.fill_(1)          # L2 flush between timed iterations (outside the event pairs)
            a.record()
            fn()
            b.record()
        barrier()
        return sum(a.elapsed_time(b) for a, b in evs)  # ms over the K steps

    sampler = ClockSampler(local)

    if args.workload == "nms":
        from oracle import boxes as OB
        n = args.nms_n
        # shard the 21 class segments over the ranks: no collective
        perm, seg = OB.class_segments(n, 21, seed=0)
        boxes = OB.sweep_boxes(n, seed=0)[perm]
        mine = F.shard_segments(21, world, rank)
        parts = [boxes[seg[s]:seg[s + 1]] for s in mine]
        local_boxes = np.ascontiguousarray(np.concatenate(parts)) if parts else np.zeros((0, 4), np.float32)
        pinned_boxes = torch.from_numpy(local_boxes).pin_memory()
        local_boxes = pinned_boxes.numpy()
        local_seg = np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.int64)
        m = F.vgg_small(F.duplo_cfg, device=local)
        bd = torch.from_numpy(local_boxes).cuda()
        l0 = m.launch_count()

        def step_dev():
            F.nms_segmented_dev(bd, local_seg, 0.25, model=m)

        picked = dict(n=0)

        def step_host():
            _, cnts = F.nms_segmented(local_boxes, local_seg, 0.25, model=m)
            picked["n"] = int(cnts.sum())

        sampler.start()
        ms = max_over_ranks(timed(step_dev, args.steps, args.warmup))
        clocks = sampler.stop()
        launches = m.launch_count() - l0
        ms_e2e = max_over_ranks(timed(step_host, args.steps, max(1, args.warmup // 2)))
        total = sum_over_ranks(float(len(local_boxes)))
        value = total * args.steps / (ms * 1e-3)
        line = dict(metric="nms_boxes_per_sec", value=value, unit="boxes/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                    config=dict(workload="nms.lua sweep: %d boxes in 21 class segments, overlap 0.25, key y2 (sort included)" % n,
                                l2="flushed between steps", sharding="class segments round-robin over ranks, no collective"),
                    clocks=clocks,
                    e2e=dict(value=total * args.steps / (ms_e2e * 1e-3), unit="boxes/s", h2d_bytes_per_step=int(local_boxes.nbytes),
                             d2h_bytes_per_step=int(8 * picked["n"] + 8 * len(mine)),
                             note="boxes from page-locked host memory; the counts and the picked indices of every class come back"),
                    gpu_launches=int(launches),
                    roofline=dict(bound="hbm", achieved=20.0 * total * args.steps / (ms * 1e-3) / 1e9 / world, peak=pk["hbm"], unit="GB/s",
                                  frac=20.0 * total * args.steps / (ms * 1e-3) / 1e9 / world / pk["hbm"], traffic=None,
                                  note="16 B/box read + 4 B class id: greedy NMS is dependency/ALU-bound, not HBM-bound (DESIGN.md)",
                                  peak_source=pk["source"]))
        if rank == 0:
            if not args.no_cpu_baseline:
                line["cpu_baseline"] = cpu_baseline_nms(n)
            print(json.dumps(line), flush=True)
        m.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    if args.workload == "train":
        # BASELINE configs[2]: vgg_small training fwd/bwd (objective.lua) 800x450, batch frames per GPU, 128 positive +
        # 128 negative anchors and 8 ground-truth boxes per frame; frames sharded over the ranks,
        # ONE all-reduce of the flat gradient (+ counters) per step
        from oracle import anchors as OA, objective as OO
        desc, cfg, params, h, w = model_setup(args)
        m = (F.vgg_small if args.model == "vgg_small" else F.vgg_large)(F.duplo_cfg if args.model == "vgg_small" else F.imgnet_cfg, device=local)
        m.load_params(params)
        oa = OA.Anchors(desc["layers"], desc["anchor_nets"], cfg["scales"])
        dims = m.output_dims(h, w)
        B = max(args.batch, 1)
        batch = []
        for s in range(B):
            pos, neg, _ = OO.synthetic_examples(oa, dims, w, h, 128, 128, 8, cfg["class_count"], seed=1000 * rank + s)
            dims_c = m.output_dims(h, w)
            pos, neg = F.clean_anchors(pos, dims_c), F.clean_anchors(neg, dims_c)
            # the example records are marshalled once by the batch iterator (BatchIterator:nextTraining's job in the
            # reference), not inside the timed lossAndGradient call
            batch.append(dict(img=OM.synthetic_frame(h, w, seed=100 * rank + s).cuda(), positive=pos, negative=neg,
                              packed=(m.pack_examples(pos), m.pack_examples(neg))))
        objective = F.create_objective(m, dist if world > 1 else None)
        l0 = m.launch_count()
        state = dict(step=0)

        def step():
            state["step"] += 1
            objective(batch, seed=state["step"])

        sampler.start()
        ms = max_over_ranks(timed(step, args.steps, args.warmup))
        clocks = sampler.stop()
        launches = (m.launch_count() - l0) // (args.steps + args.warmup)
        total = sum_over_ranks(float(B)) * args.steps
        fwd = conv_flops_per_image(desc, h, w)
        line = dict(metric="train_images_per_sec", value=total / (ms * 1e-3), unit="images/s", n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="bf16", data="synthetic",
                    config=dict(workload="%s lossAndGradient (objective.lua) %dx%d, %d frames per GPU, 128+128 anchors, 8 boxes per frame "
                                         "(BASELINE configs[2])" % (args.model, w, h, B), l2="flushed between steps",
                                sharding="frames over ranks; one all-reduce of the flat gradient per step", launches_per_step=int(launches)),
                    clocks=clocks, gpu_launches=int(launches * args.steps),
                    e2e=dict(value=total / (ms * 1e-3), unit="images/s", h2d_bytes_per_step=int(B * 256 * 96), d2h_bytes_per_step=int(B * 48),
                             note="frames resident; per frame the example lists go up and the four loss sums come back"),
                    roofline=dict(bound="tensor", kernel="conv_igemm_kernel (fwd + dgrad + wgrad launches of a step)",
                                  achieved=3 * fwd * B / (ms / args.steps * 1e-3) / 1e12, peak=pk["bf16_sustained"], unit="TFLOP/s",
                                  frac=3 * fwd * B / (ms / args.steps * 1e-3) / 1e12 / pk["bf16_sustained"], traffic=None,
                                  note="whole-step time as denominator (not the conv launches alone): lower bound of the kernel fraction"))
        if rank == 0:
            print(json.dumps(line), flush=True)
        m.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- detect workload
    desc, cfg, params, h, w = model_setup(args)
    fac, fcfg = (F.vgg_small, F.duplo_cfg) if args.model == "vgg_small" else (F.vgg_large, F.imgnet_cfg)
    m = fac(fcfg, device=local)
    m.load_params(params)
    det = F.Detector(m)
    B = args.batch
    frames = torch.stack([OM.synthetic_frame(h, w, seed=100 * rank + s) for s in range(B)])
    frames_dev = frames.cuda()
    frames_pinned = frames.pin_memory()  # e2e leg: inputs start in page-locked host memory
    frames_host = frames_pinned.numpy()
    n_det = ffi.new("int*")
    out = det._out

    def step_dev():
        rc = L.frcnn_detect_dev(m.ctx, ffi.cast("const float*", frames_dev.data_ptr()), B, h, w, out, det._cap, n_det)
        if rc != 0:
            raise RuntimeError(ffi.string(L.frcnn_last_error(m.ctx)).decode())

    def step_host():
        rc = L.frcnn_detect(m.ctx, ffi.cast("const float*", frames_host.ctypes.data), B, h, w, out, det._cap, n_det)
        if rc != 0:
            raise RuntimeError(ffi.string(L.frcnn_last_error(m.ctx)).decode())

    step_dev()
    stats = det.stats()
    # ---- synchronous steps (one frame batch, wait for its winners, L2 flushed in between): the latency figure
    ms_sync = max_over_ranks(timed(step_dev, min(args.steps, 20), args.warmup))
    sync_steps = min(args.steps, 20)
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

    def run_steps(k, ptrs, on_dev):
        busy = [False] * S
        for i in range(k):
            sl = i % S
            if busy[sl]:
                rc = L.frcnn_detect_end(ctxs[sl], out, det._cap, n_det)
                assert rc == 0, ffi.string(L.frcnn_last_error(ctxs[sl])).decode()
            rc = L.frcnn_detect_begin(ctxs[sl], ptrs[i % n_sets], on_dev, B, h, w)
            assert rc == 0, ffi.string(L.frcnn_last_error(ctxs[sl])).decode()
            busy[sl] = True
        for j in range(S):
            sl = (k + j) % S
            if busy[sl]:
                rc = L.frcnn_detect_end(ctxs[sl], out, det._cap, n_det)
                assert rc == 0, ffi.string(L.frcnn_last_error(ctxs[sl])).decode()

    def timed_region(k, warm, ptrs, on_dev):
        run_steps(max(warm, 3) * S, ptrs, on_dev)   # every context has captured and replayed its graph
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()   # legacy default stream: ordered against the (blocking) streams of the library contexts
        run_steps(k, ptrs, on_dev)
        e1.record()
        barrier()
        return e0.elapsed_time(e1)

    l0 = sum(mm.launch_count() for mm in pipe.models)
    sampler.start()
    ms = max_over_ranks(timed_region(args.steps, args.warmup, ptr_dev, 1))
    clocks = sampler.stop()
    launches = sum(mm.launch_count() for mm in pipe.models) - l0
    launches_per_step = launches // (args.steps + max(args.warmup, 3) * S)
    ms_e2e = max_over_ranks(timed_region(args.steps, args.warmup, ptr_host, 0))
    # counters (16 + 2 per frame ints) + the winners (the first 256 are copied speculatively with the counters)
    d2h = 4 * (16 + 2 * B) + max(256, int(n_det[0])) * ffi.sizeof("frcnn_detection")
    # roofline pass: the same K steps with an event pair around every conv/GEMM launch
    L.frcnn_set_profiling(m.ctx, 1)
    step_dev()
    conv_ms = conv_fl = 0.0
    conv_n = 0
    stage_ms = np.zeros(6)
    pm, pf, pn, st = ffi.new("float*"), ffi.new("double*"), ffi.new("int*"), ffi.new("float[6]")
    for _ in range(args.steps):
        flush.fill_(1)
        step_dev()
        L.frcnn_last_conv_profile(m.ctx, pm, pf, pn)
        L.frcnn_last_timings(m.ctx, st)
        conv_ms += pm[0]
        conv_fl += pf[0]
        conv_n += pn[0]
        stage_ms += np.array([st[i] for i in range(6)])
    L.frcnn_set_profiling(m.ctx, 0)
    achieved_sync = conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    # pipelined region: launches of different frames overlap, so per-launch event pairs no longer isolate the kernel;
    # the conservative figure is the algorithmic conv FLOPs of the K steps over the WHOLE timed region (every other
    # kernel of the step included in the denominator): a lower bound of the conv kernels' own rate
    achieved = (conv_fl / max(args.steps, 1)) * args.steps / (ms * 1e-3) / 1e12
    # DRAM traffic of the same launches from the committed `ncu --set full` capture (profiles/), if it covers this batch
    traffic, traffic_src = None, None
    try:
        summ = json.load(open(os.path.join(ROOT, "profiles", "r1b_summary.json")))
        rows = summ.get("full_b%d" % B)
        if rows and args.model == "vgg_small" and (h, w) == (450, 800):
            traffic = sum(r["dram_mb"] for r in rows) * 1e6
            traffic_src = "profiles/r1b_ncu_full_b%d.md (dram__bytes_read.sum + dram__bytes_write.sum over %d conv launches)" % (B, len(rows))
    except Exception:
        pass
    total_frames = sum_over_ranks(float(B)) * args.steps
    value = total_frames / (ms * 1e-3)
    peak = pk["bf16_sustained"]
    line = dict(metric="images_per_sec", value=value, unit="images/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
                config=dict(workload="%s Detector:detect %dx%d batch=%d per GPU (BASELINE configs[1])" % (args.model, w, h, B),
                            weights="seeded random init, head biases shifted so the detector stages have work (SURVEY 8d)",
                            stages=stats, in_flight=S,
                            l2="every step reads a different resident frame batch out of %d (%.0f MB > L2); no flush inside "
                               "the timed region" % (n_sets, n_sets * B * 3 * h * w * 4 / 2 ** 20),
                            sync=dict(ms_per_step=ms_sync / sync_steps, images_per_sec=B * sync_steps / (ms_sync * 1e-3),
                                      note="one step at a time (frcnn_detect_dev), L2 flushed between steps"),
                            sharding="frames over ranks, no collective", launches_per_step=int(launches_per_step)),
                clocks=clocks,
                e2e=dict(value=total_frames / (ms_e2e * 1e-3), unit="images/s", h2d_bytes_per_step=int(frames_host.nbytes),
                         d2h_bytes_per_step=int(d2h), in_flight=S,
                         note="frcnn_detect_begin on page-locked host frames (H2D inside), frcnn_detect_end reads the winners"),
                gpu_launches=int(launches_per_step * args.steps),
                roofline=dict(bound="tensor", kernel="conv_halo_kernel / conv_igemm_kernel / conv_first_kernel (all tcgen05 launches of a step)",
                              achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak if peak else None,
                              traffic=traffic, traffic_source=traffic_src,
                              how="algorithmic conv/GEMM FLOPs of the K steps / device time of the whole timed region (%d steps in "
                                  "flight: launches overlap, so the whole step is the denominator -- a lower bound of the kernels' "
                                  "own rate)" % S,
                              flops_per_step=conv_fl / max(args.steps, 1),
                              pnet_conv_flops_per_image=conv_flops_per_image(desc, h, w),
                              sync_pass=dict(achieved=achieved_sync, frac=achieved_sync / peak if peak else None,
                                             launches_per_step=conv_n // max(args.steps, 1), ms_per_step=conv_ms / max(args.steps, 1),
                                             stage_ms_per_step=dict(zip(["pnet", "decode+nms", "roi_pool", "cnet", "final_nms", "total"],
                                                                        (stage_ms / max(args.steps, 1)).round(4).tolist())),
                                             how="one step at a time on the latency schedule, an event pair around every tcgen05 launch"),
                              peak_source=pk["source"] + " bf16 sustained"))
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_detect(args, desc, cfg, params, h, w)
        print(json.dumps(line), flush=True)
    pipe.close()
    m.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
