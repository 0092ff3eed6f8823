#!/usr/bin/env python
"""Benchmark of the calibration hot path (BASELINE.json metric: AdaRound calibration iterations/s and
quantised-decode frames/s, HNeRV-3M 1280x640).

    python bench.py --gpus N --steps K --warmup W            # this repo's B200 path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores

A step is ONE AdaRound (phase-2) calibration iteration -- forward, loss, backward, rounding
regulariser, Adam -- on a batch of 2 synthetic 1280x640 frames per GPU (frame-sharded data parallel:
global batch 2*N, one NCCL all-reduce of the flat dW/db buffer per step).  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "adaround_calib_iters_per_s"
UNIT = "it/s (batch-2 iterations; frame-sharded DP processes N of them per step)"
DEFAULT_BITS = [6, 5, 4, 5, 5, 6, 6]
HYPER = dict(weight=0.01, b_range=(20, 2), warmup=0.2, p=2.0, lr=0.003, iters=21000)
TRAFFIC_FILE = "r02x_traffic.json"   # dram bytes per launch of the dominant kernels, from the committed `ncu --set full` capture
MMA_PASSES = {"conv_fwd": 3, "conv_dgrad": 3, "conv_wgrad": 3, "head_fwd_loss": 3}  # bf16 MMAs per fp32-equivalent product


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="hnerv-bunny-3m")
    ap.add_argument("--batch", type=int, default=2, help="frames per GPU per iteration")
    ap.add_argument("--precision", type=int, nargs="+", default=DEFAULT_BITS)
    ap.add_argument("--hadamard", action="store_true")
    ap.add_argument("--frames", type=int, default=8, help="synthetic frames resident per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--decode-steps", type=int, default=10)
    ap.add_argument("--artefact", action="store_true", help="--mode decode: decode from the packed artefact written by this run")
    ap.add_argument("--stage", type=int, default=5, help="--mode block: decoder stage (block) to reconstruct")
    ap.add_argument("--no-hadamard-record", action="store_true", help="skip the --hadamard sub-record of the calibration line")
    ap.add_argument("--mode", default="calib", choices=["calib", "decode", "block", "train", "omega"],
                    help="decode: quantised-decode throughput only (any workload, e.g. hnerv-1080p-12m)")
    return ap.parse_args()


def config_of(args, world):
    return {"workload": f"{args.workload} 1280x640 AdaRound phase-2 iteration, W-mixed {' '.join(map(str, args.precision))}"
                        if "bunny" in args.workload else args.workload,
            "frames_per_gpu_per_step": args.batch, "global_batch": args.batch * world, "hadamard": bool(args.hadamard),
            "parallelism": f"dp{world} (frame-sharded, NCCL all-reduce of dW)" if world > 1 else "single GPU",
            "l2": "no explicit flush: each step streams >2 GB of activations, far above the 126 MB L2",
            "launch": "CUDA graph replay of the iteration; when N > 1 the NCCL all-reduce of dW is a node of the same graph",
            **{k: v for k, v in HYPER.items()}}


# ------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(index), "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 6 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's OWN code on the host cores (oracle/_ref, staged by oracle/make_ref.py),
# or -- when that is absent -- the oracle's restatement of calib_model.py.  Neither imports neuroquant_b200.
# ------------------------------------------------------------------------------------------------------
def cpu_iteration_rate(args, steps, warmup):
    """(iterations/s, decode frames/s, cores, kind, note) of the reference's CPU path on this box for args.workload."""
    from oracle import ref_runner

    cores = os.cpu_count() or 1
    if ref_runner.available() and os.environ.get("NQ_CPU_BASELINE", "reference") != "port":
        per_iter, dec, _ = ref_runner.time_calibration(args.workload, args.precision, args.hadamard, args.batch, steps, warmup,
                                                       HYPER, cores)
        return 1.0 / per_iter, args.batch / dec, cores, "reference", \
            "the unmodified reference (oracle/_ref) through quantization.model_reconstruction / QuantModel.forward"
    step, decode, cores = oracle_iteration_fn(args)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    decode()
    td = time.perf_counter()
    decode()
    td = time.perf_counter() - td
    return steps / dt, args.batch / td, cores, "port", "oracle port of calib_model.py (oracle/_ref not staged)"


def oracle_iteration_fn(args):
    """Returns (step_fn, decode_fn, cores): step_fn() runs one AdaRound iteration of the reference
    algorithm (oracle/nq_oracle.py, pinned to the reference by tests/golden) on CPU."""
    from oracle import nq_oracle as O
    from oracle import workloads as W

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    arch, cfg = W.WORKLOADS[args.workload]
    qd = O.QuantDecoder(W.random_stages(cfg, arch, 903), args.precision, args.hadamard)
    qd.start_adaround()
    alphas = []
    for q in qd.q:
        q.alpha_w.requires_grad_(True)
        q.alpha_b.requires_grad_(True)
        alphas += [q.alpha_w, q.alpha_b]
    opt = torch.optim.Adam(alphas, lr=HYPER["lr"])
    gen = torch.Generator().manual_seed(903)
    c, h, w = W.embed_shape(cfg, arch)
    embed = torch.randn(args.batch, c, h, w, generator=gen)
    frames = torch.rand(args.batch, 3, cfg["crop_h"], cfg["crop_w"], generator=gen)

    def step():
        out = qd.forward(embed)
        opt.zero_grad()
        rec = O.lp_loss(out, frames, p=HYPER["p"])
        rnd = sum(HYPER["weight"] * O.round_reg(q.alpha_w, 10.0) for q in qd.q)
        (rec + rnd).backward()
        opt.step()
        return float(rec.detach())

    def decode():
        with torch.no_grad():
            return qd.forward(embed)

    return step, decode, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.mode == "omega":
        return run_omega_reference(args)
    warm = min(args.warmup, 2)  # a CPU iteration is ~0.6 s; two warm-ups page everything in
    its, dec_fps, cores, kind, note = cpu_iteration_rate(args, args.steps, warm)
    sample = f"{args.steps} AdaRound iterations after {warm} warm-up, batch {args.batch}, {args.workload}: {note}, {cores} threads"
    if args.mode == "decode":
        metric, unit, val = "quantized_decode_frames_per_s", "frames/s", dec_fps
    else:
        metric, unit, val = METRIC, UNIT, its
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 / its, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_of(args, 1),
        "cpu_baseline": {"value": val, "unit": unit, "cores": cores, "kind": kind, "sample": sample},
        "decode": {"metric": "quantized_decode_frames_per_s", "value": dec_fps, "unit": "frames/s", "batch": args.batch},
        "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------
class HostLoader(list):
    """The `gt` loader of the end-to-end leg: a sequence of sample dicts whose frames sit in pinned HOST memory as uint8
    (what VideoDataSet(as_uint8=True) + DataLoader(pin_memory=True) hand to model_reconstruction)."""
    batch_size = None


def build_quant_model(args, cfg, arch, params, hadamard):
    """QuantModel of the workload with the seeded decoder weights, bit-widths set and scales initialised, exactly as
    methods/calibrate_network.py:218-238 prepares it for model_reconstruction."""
    from neuroquant_b200.models import HNeRV, NeRV
    from neuroquant_b200.quantization import QuantModel

    torch.manual_seed(1)
    model = (HNeRV if arch == "hnerv" else NeRV)(dict(cfg))
    convs = [model.decoder[0]] + [blk.conv[0] for blk in list(model.decoder)[1:]] + [model.head_layer]
    with torch.no_grad():
        for conv, (w, b) in zip(convs, params):
            conv.weight.copy_(w)
            conv.bias.copy_(b)
    qnn = QuantModel(model.cuda(), hadamard=hadamard,
                     weight_quant_params={"n_bits": 8, "channel_wise": True, "scale_method": "max"}).cuda()
    qnn.set_bitwidth(list(args.precision))
    qnn.eval()
    qnn.set_quant_state(True)
    return qnn


def run_b200(args):
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback; use --impl reference for the host baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import neuroquant_b200 as nq
    from neuroquant_b200.engine import AdamState
    from neuroquant_b200.workloads import WORKLOADS, conv_flops, embed_shape, random_decoder

    arch, cfg = WORKLOADS[args.workload]
    geoms, params = random_decoder(cfg, arch, 903)
    if args.mode == "omega":
        return run_omega(args, cfg, arch, geoms, params, world, rank)
    stages = [nq.QuantStage(g, w.cuda(), b.cuda(), nb, args.hadamard) for g, (w, b), nb in zip(geoms, params, args.precision)]
    eng = nq.DecoderEngine(stages)
    eng.init_scales()
    eng.start_adaround()
    c, h0, w0 = embed_shape(cfg, arch)
    if args.mode == "decode":
        return run_decode_only(args, eng, cfg, arch, geoms, world, rank, local)
    if args.mode == "block":
        return run_block(args, eng, cfg, arch, geoms, world, rank)
    if args.mode == "train":
        del eng
        return run_train(args, cfg, arch, world, rank)
    H, W = cfg["crop_h"], cfg["crop_w"]
    gen = torch.Generator().manual_seed(903 + rank)
    F = max(args.frames, args.batch)
    embeds_h = torch.randn(F, c, h0, w0, generator=gen).pin_memory()
    frames_u8_h = torch.randint(0, 256, (F, 3, H, W), generator=gen, dtype=torch.uint8).pin_memory()
    embeds_d = embeds_h.cuda()
    frames_d = frames_u8_h.cuda()  # uint8, as the calibration loader hands them out (value / 255 inside the head kernel)
    B = args.batch
    mean_pixels = float(B * world * H * W)
    reg_w, reg_b = HYPER["weight"], 10.0  # mid-schedule temperature: regulariser on, as in 80% of the run

    from neuroquant_b200.calibration import GraphedStep
    use_graph = os.environ.get("NQ_GRAPH", "1") != "0"

    def make_stepper(eng_):
        opt_ = AdamState([t_ for s in eng_.stages for t_ in (s.alpha_w, s.alpha_b)], lr=HYPER["lr"])
        graphed = {}
        capture = world == 1 or os.environ.get("NQ_GRAPH_DP", "1") != "0"

        def step(i, embed, frames, eager=False):
            """One AdaRound iteration exactly as CalibrationLoop.iteration runs it."""
            if use_graph and not eager:
                if "g" not in graphed:
                    graphed["g"] = GraphedStep(eng_, opt_, embed, frames, HYPER["p"], mean_pixels, None, world, capture=capture)
                graphed["g"].run(embed, frames, reg_w, reg_b)
                return
            eng_.forward(embed, train=True, target=frames, p_norm=HYPER["p"], mean_pixels=mean_pixels, want_img=False)
            flat = eng_.backward()
            if world > 1:
                dist.all_reduce(flat)
            grads = eng_.param_grads(1.0, reg_w, reg_b)
            opt_.step([g for pair in grads for g in pair])
            eng_.launches += len(opt_.params)
            eng_.invalidate()
        return step

    step = make_stepper(eng)

    def resident(i):
        o = (i * B) % (F - B + 1)
        return embeds_d[o:o + B], frames_d[o:o + B]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms), t0, t1

    # ---- device-resident throughput
    for i in range(args.warmup):
        step(i, *resident(i))
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = eng.launches
    ms, t0, t1 = timed(lambda i: step(i, *resident(i)), args.steps)
    launches = eng.launches - l0
    clocks = sampler.stop(t0, t1) if sampler else None
    value = args.steps * world / (ms * 1e-3)

    # ---- per-kernel timing of the convolution launches (CUDA events on the launch stream)
    kern = eng.kernel_profile(lambda: step(0, *resident(0), eager=True), reps=5)

    # ---- quantised decode (hard rounding, weights static -> packed once, Q9)
    eng.soft_w = False
    eng.invalidate()
    eng.forward(embeds_d[:B])
    for i in range(3):
        eng.forward(embeds_d[:B], reuse_weights=True)
    ms_d, _, _ = timed(lambda i: eng.forward(resident(i)[0], reuse_weights=True), args.decode_steps)
    decode_fps = args.decode_steps * B * world / (ms_d * 1e-3)
    kern_dec = eng.kernel_profile(lambda: eng.forward(embeds_d[:B], reuse_weights=True), reps=5)
    eng.soft_w = True
    eng.invalidate()

    # ---- the same step with --hadamard (BASELINE configs[2]) as a sub-record, so that the scaling runs cover it
    had = None
    if not args.hadamard and not args.no_hadamard_record:
        st_h = [nq.QuantStage(g, w.cuda(), b.cuda(), nb, True) for g, (w, b), nb in zip(geoms, params, args.precision)]
        eng_h = nq.DecoderEngine(st_h)
        eng_h.init_scales()
        eng_h.start_adaround()
        step_h = make_stepper(eng_h)
        for i in range(max(args.warmup, 3)):
            step_h(i, *resident(i))
        n_h = max(5, args.steps // 2)
        ms_h, _, _ = timed(lambda i: step_h(i, *resident(i)), n_h)
        had = {"value": n_h * world / (ms_h * 1e-3), "unit": UNIT, "ms_per_step": ms_h / n_h, "steps": n_h,
               "config": "same step with --hadamard (weights rotated along C_in, BASELINE configs[2])"}
        del eng_h, step_h, st_h
        torch.cuda.empty_cache()

    # ---- end to end THROUGH THE PLUG-IN CALL: quantization.model_reconstruction(qnn, cali_data, gt=<host loader>) as
    # methods/calibrate_network.py makes it.  Frames start in pinned HOST memory every step (uint8, frame_residency
    # 'stream': nothing is cached in HBM), the embeddings too; the loss goes back to the host every step.  A loader of
    # W + K + 1 mini-batches with iters = W + K + 1 gives 0 step-size epochs and one AdaRound epoch (calib_model.py:144,
    # 203-206); CUDA events bracket iterations W+1 .. W+K from the per-iteration callback.
    del eng, step
    torch.cuda.empty_cache()
    from neuroquant_b200.quantization import model_reconstruction
    import neuroquant_b200.quantization.calib_model as cm
    qnn = build_quant_model(args, cfg, arch, params, args.hadamard)
    qnn(embeds_d[:B])  # first quantised forward: initialises the step sizes (calibrate_network.py:235-238)
    We, Ke = max(args.warmup, 3), args.steps
    GB = B * world
    Fg = max(F, GB)
    if Fg > F:  # the global batch needs more distinct frames than one rank holds
        gen2 = torch.Generator().manual_seed(905)
        embeds_h = torch.randn(Fg, c, h0, w0, generator=gen2).pin_memory()
        frames_u8_h = torch.randint(0, 256, (Fg, 3, H, W), generator=gen2, dtype=torch.uint8).pin_memory()
    loader = HostLoader()
    loader.batch_size = GB
    for i in range(We + Ke + 1):
        o = (i * GB) % (Fg - GB + 1)
        loader.append({"img": frames_u8_h[o:o + GB], "idx": torch.arange(o, o + GB), "norm_idx": torch.arange(o, o + GB).float() / Fg})
    loss_ring = torch.zeros(8).pin_memory()
    ev = {}

    def on_iteration(phase, count, loss):
        loss_ring[count % 8:count % 8 + 1].copy_(loss.view(1), non_blocking=True)  # device -> host, every step
        if count == We:
            sync_all()
            ev["h2d0"] = cm.LAST_RUN.get("_src").h2d_bytes if cm.LAST_RUN.get("_src") else 0
            ev["t0"] = torch.cuda.Event(enable_timing=True)
            ev["t0"].record()
        elif count == We + Ke:
            ev["t1"] = torch.cuda.Event(enable_timing=True)
            ev["t1"].record()
            ev["h2d1"] = cm.LAST_RUN.get("_src").h2d_bytes if cm.LAST_RUN.get("_src") else 0

    model_reconstruction(qnn, cali_data=embeds_h, gt=loader, arch=arch, batch_size=GB, iters=We + Ke + 1, weight=HYPER["weight"],
                         opt_mode="mse", hadamard=args.hadamard, b_range=HYPER["b_range"], warmup=0.0, p=HYPER["p"], lr=HYPER["lr"],
                         frame_residency="stream", on_iteration=on_iteration)
    torch.cuda.synchronize()
    ms_e = torch.tensor([ev["t0"].elapsed_time(ev["t1"])], device="cuda")
    if world > 1:
        dist.all_reduce(ms_e, op=dist.ReduceOp.MAX)
    ms_e = float(ms_e)
    e2e_val = Ke * world / (ms_e * 1e-3)
    h2d = (ev["h2d1"] - ev["h2d0"]) / Ke if ev.get("h2d1") else B * (c * h0 * w0 * 4 + 3 * H * W)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    flops_iter = 3.0 * conv_flops(geoms, h0, w0, B)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    # a run of a few dozen steps ends before the power controller reacts (burst clocks): the burst peak is the right
    # denominator; a long run (>= 1 s timed) sits under the power cap: the sustained peak
    long_run = ms >= 1000.0
    peak_key = "bf16_tflops_sustained" if long_run else "bf16_tflops"
    peak_tf = peaks.get(peak_key, 1400.0 if long_run else 1590.0)
    peak_src = (f"MEASURED_PEAKS.json {peak_key} (of measured)" if peaks else "B200_PROFILING.md fallback (of fallback)") + \
        (": timed region >= 1 s, under the power cap" if long_run else ": timed region < 1 s, burst clocks")
    top = kern[0]
    conv_ms = sum(k["ms"] for k in kern)
    traffic = None
    try:  # dram__bytes_read + dram__bytes_write of this launch from the committed `ncu --set full` capture
        tr = json.load(open(os.path.join(ROOT, "profiles", TRAFFIC_FILE)))
        if args.workload == "hnerv-bunny-3m" and B == 2:
            traffic = tr.get(top["kernel"])
    except OSError:
        pass
    passes = MMA_PASSES.get(top["kernel"].split("[")[0], 3)
    roof = {"bound": "tensor", "kernel": top["kernel"], "achieved": top["tflops"], "peak": peak_tf, "unit": "TFLOP/s",
            "frac": top["tflops"] / peak_tf, "traffic": traffic, "ms_per_launch": top["ms"], "flops_per_launch": top["flops"],
            "note": "achieved = algorithmic 2*M*N*K of the fp32-equivalent convolution / CUDA-event time; the kernel issues "
                    f"{passes} bf16 MMAs per product (split hi/lo operands), so its tensor-pipe rate is {passes}x this figure",
            "tensor_pipe_frac": passes * top["tflops"] / peak_tf,
            "share_of_step": top["ms"] / (ms / args.steps),
            "peak_source": peak_src,
            "conv_kernels_ms": {k["kernel"]: round(k["ms"], 4) for k in kern}, "conv_share_of_step": conv_ms / (ms / args.steps),
            "step_tflops": flops_iter / (ms / args.steps * 1e-3) / 1e12, "step_frac": flops_iter / (ms / args.steps * 1e-3) / 1e12 / peak_tf}
    gf_frame = conv_flops(geoms, h0, w0, 1) / 1e9
    dtop = kern_dec[0]
    dec = {"metric": "quantized_decode_frames_per_s", "value": decode_fps, "unit": "frames/s", "batch": B,
           "ms_per_batch": ms_d / args.decode_steps, "steps": args.decode_steps, "gflop_per_frame": gf_frame,
           "roofline": {"bound": "tensor", "kernel": dtop["kernel"], "achieved": dtop["tflops"], "peak": peak_tf, "unit": "TFLOP/s",
                        "frac": dtop["tflops"] / peak_tf, "ms_per_launch": dtop["ms"],
                        "note": "hard-rounded weights are integers (one exact bf16 plane), activations split hi/lo: 2 MMAs per product",
                        "tensor_pipe_frac": 2 * dtop["tflops"] / peak_tf,
                        "whole_decode_tflops": decode_fps / world * gf_frame / 1e3,
                        "whole_decode_frac": decode_fps / world * gf_frame / 1e3 / peak_tf,
                        "conv_kernels_ms": {k["kernel"]: round(k["ms"], 4) for k in kern_dec}}}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16x2 split operands, fp32 accumulate (tcgen05)" if os.environ.get("NQ_CONV", "tc") != "simt" else "f32",
        "data": "synthetic", "config": config_of(args, world),
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e / Ke, "steps": Ke,
                "path": "quantization.model_reconstruction(qnn, cali_data=<pinned host>, gt=<pinned host uint8 loader>, "
                        "frame_residency='stream'): H2D of every batch on a side stream one batch ahead, loss read back every step"},
        "decode": dec,
        "tflops_effective": flops_iter / (ms / args.steps * 1e-3) / 1e12,
        "roofline": roof,
    }
    if had is not None:
        out["hadamard"] = had
    if world == 1 and not args.no_cpu_baseline:
        its, dec_fps, cores, kind, note = cpu_iteration_rate(args, 3, 1)
        out["cpu_baseline"] = {"value": its, "unit": UNIT, "cores": cores, "kind": kind,
                               "sample": f"3 AdaRound iterations (batch {B}) after 1 warm-up, same workload: {note}; "
                                         f"decode {dec_fps:.2f} frames/s over two batch-{B} decodes"}
        out["decode"]["cpu_baseline"] = {"value": dec_fps, "unit": "frames/s", "cores": cores, "kind": kind,
                                         "sample": f"two hard-rounded batch-{B} decodes after one warm-up"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_omega(args, cfg, arch, geoms, params, world, rank):
    """BASELINE.json configs[3]: bit_assign's Omega search.  A step = one search over ALL 7^L per-layer bit-width
    configurations (bits 2..8) under an average-bit budget: the perturbation of every layer at every bit-width
    (QuantModel + get_perturbation), the Gram table of Omega from 49 + 1029 forward jets over 10 mini-batches of 2 frames
    (farmed over the ranks, one all-reduce of 1078 doubles), and the table search kernel.  metric: candidate
    configurations scored per second, whole job; the reference scores ONE candidate per Hessian-vector product pass."""
    import torch.distributed as dist
    from neuroquant_b200.methods.bit_assign import search_bit_assignment
    from neuroquant_b200.workloads import embed_shape

    H, W = cfg["crop_h"], cfg["crop_w"]
    c, h0, w0 = embed_shape(cfg, arch)
    from neuroquant_b200.models import HNeRV, NeRV
    torch.manual_seed(1)
    fp = (HNeRV if arch == "hnerv" else NeRV)(dict(cfg)).cuda()
    convs = [fp.decoder[0]] + [blk.conv[0] for blk in list(fp.decoder)[1:]] + [fp.head_layer]
    with torch.no_grad():
        for conv, (w, b) in zip(convs, params):
            conv.weight.copy_(w.cuda())
            conv.bias.copy_(b.cuda())
    gen = torch.Generator().manual_seed(903)
    n_b, B = 10, args.batch
    frames = torch.rand(n_b * B, 3, H, W, generator=gen).cuda()
    embeds = torch.randn(n_b * B, c, h0, w0, generator=gen).cuda()
    # the loader hands out frames; the embeddings of a random-init encoder carry no information, so the decoder inputs
    # are the seeded embeddings above (as in the calibration benchmark)
    loader = [{"img": frames[i:i + B], "norm_idx": torch.arange(i, i + B).float().cuda() / (n_b * B), "idx": torch.arange(i, i + B)}
              for i in range(0, n_b * B, B)]
    lookup = {int(s["idx"][0]): embeds[int(s["idx"][0]):int(s["idx"][0]) + B] for s in loader}
    state = {"i": 0}

    def encode(x):
        k = (state["i"] % n_b) * B
        state["i"] += 1
        return lookup[k]
    fp.encode = encode
    options = [2, 3, 4, 5, 6, 7, 8]
    budget = 4.8

    def one():
        state["i"] = 0
        return search_bit_assignment(arch, fp, loader, embeds, options, budget, batch_size=B)

    if world > 1:
        dist.barrier()
    for _ in range(min(args.warmup, 1)):
        one()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0"))) if rank == 0 else None
    t0 = time.perf_counter()
    e0.record()
    steps = max(1, min(args.steps, 3))
    for _ in range(steps):
        bits, score, avg_bits, table = one()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms) / steps
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    n_cfg = len(options) ** table.L
    out = {"metric": "omega_candidates_per_s", "value": n_cfg / (ms * 1e-3), "unit": "bit-width configurations scored / s",
           "n_gpus": world, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "bf16x2 split operands, fp32 accumulate (tcgen05); table in f64",
           "data": "synthetic",
           "config": {"workload": f"{args.workload} bit_assign Omega search: {table.L} layers x bits {options}, average bits <= {budget}, "
                                  f"{n_b} mini-batches of {B} frames", "configurations": n_cfg, "forward_jets": len(table.directions()),
                      "parallelism": f"jets farmed over {world} GPU(s), one all-reduce of {len(table.directions())} doubles"},
           "clocks": sampler.stop(t0, t1) if sampler else None, "gpu_launches": table.eng.launches,
           "best": {"bits": bits, "omega": score, "avg_bits": avg_bits},
           "reference_candidates": {"[6,5,4,5,5,6,6]": table.score([6, 5, 4, 5, 5, 6, 6]), "[2,3,4,6,4,4,2]": table.score([2, 3, 4, 6, 4, 4, 2])},
           "seconds_per_jet": ms * 1e-3 / len(table.directions()) * world,
           "e2e": {"value": n_cfg / (ms * 1e-3), "unit": "bit-width configurations scored / s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 8 * len(table.directions()),
                   "note": "the search starts from HBM-resident frames (20 frames); the table travels to the host once per search"}}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_omega_reference(args):
    """The reference's own sensitivity_criterion('omega') (bit_assign.py:171-203: double-backward Hessian-vector product)
    on the host cores, bounded: ONE candidate over `n_b` mini-batches of 2 frames instead of 10 (a full candidate takes
    minutes on CPU); candidates/s is scaled to the full 10 batches."""
    import copy
    import importlib.util
    from oracle import ref_runner
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if not ref_runner.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not staged: the Omega arm needs the reference's bit_assign.py"}))
        return
    from oracle import workloads as W
    qnn, arch, cfg, _ = ref_runner.build(args.workload, [6, 5, 4, 5, 5, 6, 6], False)
    spec = importlib.util.spec_from_file_location("ref_bit_assign", os.path.join(ref_runner.REF, "methods", "bit_assign.py"))
    ba = importlib.util.module_from_spec(spec)
    _cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self   # bit_assign.py:110 hard-codes .cuda(); this arm is the CPU path
    try:
        spec.loader.exec_module(ba)
        gen = torch.Generator().manual_seed(903)
        c, h, w = W.embed_shape(cfg, arch)
        n_b = max(1, min(args.steps, 2))
        frames = torch.rand(n_b * 2, 3, cfg["crop_h"], cfg["crop_w"], generator=gen)
        embeds = torch.randn(n_b * 2, c, h, w, generator=gen)
        with torch.no_grad():
            qnn(embeds[:2])
        net = ref_runner.build.fp_model   # the full-precision network (bit_assign.py:364 passes a copy of it)
        loader = [{"img": frames[i:i + 2], "norm_idx": torch.arange(i, i + 2).float() / (2 * n_b), "idx": torch.arange(i, i + 2)}
                  for i in range(0, 2 * n_b, 2)]
        if arch == "hnerv":  # decoder inputs = the seeded embeddings (the random-init encoder is not part of the path timed)
            calls = {"n": 0}

            def encode(x):
                k = 2 * (calls["n"] % n_b)
                calls["n"] += 1
                return embeds[k:k + 2]
            net.encode = encode
        t0 = time.perf_counter()
        om = ba.sensitivity_criterion("omega", arch, net, qnn, loader, use_cuda=False)
        dt = time.perf_counter() - t0
    finally:
        torch.Tensor.cuda = _cuda
    per_candidate = dt / n_b * 10
    val = 1.0 / per_candidate
    unit = "bit-width configurations scored / s"
    print(json.dumps({"impl": "reference", "metric": "omega_candidates_per_s", "value": val, "unit": unit, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * per_candidate, "higher_is_better": True,
                      "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": f"{args.workload} bit_assign Omega: one candidate = Hessian-vector product over 10 mini-batches of 2 frames"},
                      "cpu_baseline": {"value": val, "unit": unit, "cores": cores, "kind": "reference",
                                       "sample": f"the reference's sensitivity_criterion('omega') over {n_b} of the 10 mini-batches "
                                                 f"({dt:.1f} s), scaled to 10; omega = {float(om):.3e}"},
                      "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def run_decode_only(args, eng, cfg, arch, geoms, world, rank, local):
    """Quantised decode (hard AdaRound rounding, weights packed once -- SURVEY Q9) of synthetic embeddings; frames are
    sharded across ranks with no collective on the data path."""
    import torch.distributed as dist
    from neuroquant_b200.workloads import conv_flops, embed_shape

    c, h0, w0 = embed_shape(cfg, arch)
    B = args.batch
    gen = torch.Generator().manual_seed(903 + rank)
    embeds = torch.randn(max(args.frames, B), c, h0, w0, generator=gen).cuda()
    eng.soft_w = False
    eng.invalidate()
    eng.forward(embeds[:B])
    artefact_bytes = None
    if args.artefact:  # write the packed artefact of this decoder, drop the live model, decode from the file
        import tempfile
        from neuroquant_b200.artefact import PackedDecoder, save_artefact
        path = os.path.join(tempfile.mkdtemp(), f"{args.workload}.nqb")
        artefact_bytes = save_artefact(eng, path)
        want = eng.forward(embeds[:B], reuse_weights=True).clone()
        eng = PackedDecoder(path).engine
        assert torch.equal(eng.forward(embeds[:B]), want), "decode from the artefact differs from the live model"
    for _ in range(max(args.warmup, 3)):
        eng.forward(embeds[:B], reuse_weights=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng.launches
    e0.record()
    for i in range(args.steps):
        o = (i * B) % (embeds.shape[0] - B + 1)
        eng.forward(embeds[o:o + B], reuse_weights=True)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.barrier()
    ms = float(ms)
    if rank == 0:
        fps = args.steps * B * world / (ms * 1e-3)
        gf = conv_flops(geoms, h0, w0, 1) / 1e9
        print(json.dumps({"metric": "quantized_decode_frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": world,
                          "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": eng.dtype_name,
                          "data": "synthetic", "config": {"workload": args.workload + " quantised decode", "batch": B},
                          "gpu_launches": eng.launches - l0, "gflop_per_frame": gf,
                          "tflops_effective": fps * gf / 1e3,
                          **({"artefact_bytes": artefact_bytes, "artefact": "decoded from the packed artefact (bit-identical to the live model)"}
                             if artefact_bytes else {})}))
    if world > 1:
        dist.destroy_process_group()


def run_block(args, eng, cfg, arch, geoms, world, rank):
    """Block-wise reconstruction (calib_block.py): iterations/s of ONE decoder block learning its
    rounding against cached full-precision outputs -- per iteration a random mini-batch gather from the HBM-resident
    cache, QDrop mixing (input_prob 0.5), then quantization/calib_block.BlockStep (soft fake-quant, tcgen05 forward, fused
    loss + activation backward, tcgen05 weight gradient, fused Jacobian + Adam).  Single GPU: blocks are calibrated one
    after the other and a block has no exchange step."""
    from neuroquant_b200.quantization.calib_block import BlockStep
    from neuroquant_b200.workloads import embed_shape

    if world > 1:
        if rank == 0:
            print(json.dumps({"metric": "block_recon_iters_per_s", "unavailable": "block-wise reconstruction is a single-GPU loop (replicas only)"}))
        return
    k = args.stage
    if not 1 <= k < len(geoms) - 1:
        raise SystemExit("--stage must be a decoder block (1 .. n_stages - 2)")
    _, h, w = embed_shape(cfg, arch)
    for g in geoms[:k]:
        h, w = h * g.rh, w * g.rw
    g = geoms[k]
    st = eng.stages[k]
    B, F = args.batch, max(args.frames, args.batch)
    gen = torch.Generator().manual_seed(903)
    inp = torch.randn(F, g.cin, h, w, generator=gen).cuda()
    sym = inp + 0.01 * torch.randn(F, g.cin, h, w, generator=gen).cuda()
    out = torch.randn(F, g.c_grp, h * g.rh, w * g.rw, generator=gen).cuda() * 0.3
    from neuroquant_b200.quantization.calib_block import cache_to_engine_layout
    step = BlockStep(st, B, h, w, lr=HYPER["lr"])
    inp_s, sym_s, out_c = cache_to_engine_layout(step, inp, sym, out)

    from neuroquant_b200.quantization.calib_block import assemble_batch
    cur = torch.empty_like(inp_s[:, :B])
    idx = torch.zeros(B, dtype=torch.int32, device="cuda")
    use_graph = os.environ.get("NQ_GRAPH", "1") != "0"

    def it(i):  # exactly the loop body of quantization/calib_block.block_reconstruction
        idx.copy_(torch.randperm(F)[:B])
        assemble_batch(inp_s, sym_s, idx, 0.5, cur)
        step.run_cached(cur, out_c, idx, HYPER["weight"], 10.0, HYPER["p"], graph=use_graph)

    for i in range(max(args.warmup, 3)):
        it(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        it(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    flops = 2.0 * B * h * w * g.cout * g.cin * g.k * g.k * 2  # forward + weight gradient
    print(json.dumps({"metric": "block_recon_iters_per_s", "value": 1e3 / ms, "unit": f"it/s (batch-{B} iterations of one block)",
                      "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": eng.dtype_name, "data": "synthetic",
                      "config": {"workload": f"{args.workload} block {k}: {g.cin}->{g.cout} k{g.k} at {h}x{w}, batch {B}, QDrop 0.5",
                                 "cache_frames": F},
                      "gpu_launches": 9 * args.steps, "tflops_algorithmic": flops / (ms * 1e-3) / 1e12}))


def run_train(args, cfg, arch, world, rank):
    """FP32 regression training (methods/regress.py): iterations/s of the reference's training step --
    adjust_lr, forward, 'l2' loss, backward, Adam on every parameter -- with the decoder on the engine's tcgen05 kernels
    in full-precision mode (methods/regress.DecoderTrainer) and, for HNeRV, the stock-PyTorch ConvNeXt encoder in front.
    Frames resident in HBM; single GPU as in the reference."""
    from types import SimpleNamespace
    from neuroquant_b200.methods.regress import DecoderTrainer
    from neuroquant_b200.models import HNeRV, NeRV
    from neuroquant_b200.utils import adjust_lr

    if world > 1:
        if rank == 0:
            print(json.dumps({"metric": "fp_train_iters_per_s", "unavailable": "regression training is single-GPU, as in the reference"}))
        return
    torch.manual_seed(903)
    cfg = dict(cfg)
    cfg.setdefault("diff_enc", False)
    model = (HNeRV if arch == "hnerv" else NeRV)(cfg).cuda().train()
    B, F = args.batch, max(args.frames, args.batch)
    gen = torch.Generator().manual_seed(903)
    frames = torch.rand(F, 3, cfg["crop_h"], cfg["crop_w"], generator=gen).cuda()
    norm_idx = (torch.arange(F).float() / F).cuda()
    targs = SimpleNamespace(lr=float(cfg.get("learning_rate", 1e-3)), lr_type="cosine_0.1_1_0.1")
    trainer = DecoderTrainer(model, arch, targs.lr)
    total = max(args.warmup, 3) + args.steps

    def it(i):
        adjust_lr(trainer, i / total, targs)
        idx = torch.arange(i * B, i * B + B, device="cuda") % F
        img = frames[idx]
        return trainer.step(img if arch == "hnerv" else norm_idx[idx], img)

    for i in range(max(args.warmup, 3)):
        first, _ = it(i)
    torch.cuda.synchronize()
    l0 = trainer.launches + trainer.runner.engine.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        last, _ = it(max(args.warmup, 3) + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    eng = trainer.runner.engine
    print(json.dumps({"metric": "fp_train_iters_per_s", "value": 1e3 / ms, "unit": f"it/s (batch-{B} training iterations)",
                      "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": eng.dtype_name, "data": "synthetic",
                      "config": {"workload": f"{args.workload} FP32 regression step, batch {B}, loss l2, Adam, cosine lr",
                                 "encoder": "stock PyTorch ConvNeXt (autograd), fed by the engine's d_embed" if arch == "hnerv" else "none",
                                 "l2_inputs": "frames resident in HBM"},
                      "gpu_launches": (eng.launches + trainer.launches - l0), "loss_first": float(first), "loss_last": float(last)}))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
