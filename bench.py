#!/usr/bin/env python
"""bench.py — headline benchmark of the captioning hot path (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload train|encoder]

Headline workload = the configuration BASELINE.json's metric ("train images/sec ... at 1/2/4/8 B200") is quoted on,
configs[3]: the trainMultiGPU.py step — Encoder fine-tuned from startingLayer=7 + LSTM-attention decoder, teacher
forcing, bf16 compute, batch 32 per GPU, 256x256 synthetic images, 52-token synthetic captions, packed cross-entropy
+ attention regulariser, clamp +-5, Adam.  One "step" = one such train step.  N > 1 (launched by torchrun, one rank
per GPU): DistributedDataParallel over NCCL, disjoint per-rank batches (weak scaling), gradient all-reduce
overlapped with the explicit backward.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      dominant kernel of the step (the tcgen05 GEMM, ~50 % of the kernel time): algorithmic FLOPs /
                CUDA-event time summed over its launches, live, vs MEASURED_PEAKS.json
  kernels       per-kernel-kind breakdown from the same instrumented pass
  cpu_baseline  the oracle (CPU restatement of the reference's step) on the host cores, bounded sample
  e2e           same metric through the public call with pinned HOST inputs and a D2H read of the loss every step
  extra         the other BASELINE.json configs: encoder_forward (configs[1], with its own roofline / e2e),
                Transformer train step (configs[2]), batched beam search (configs[4]), torch-eager encoder on B200
``--workload encoder`` makes configs[1] (Encoder.forward, batch 64) the headline line instead.
``--impl reference`` times the reference's CPU implementation of the same path (the oracle port: the reference
itself is pure Python on torchvision and /root/reference does not exist on the GPU box) with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENCODER_GFLOP_PER_IMAGE = 40.11   # 20.054 GMAC per 256x256 image
TRAIN_BATCH = 32                  # BASELINE.json configs[3]: per-GPU batch of the trainMultiGPU.py step
V = 9490
WORDMAP = {"<pad>": 0, "<unk>": V - 3, "<start>": V - 2, "<end>": V - 1}
L2_BYTES = 126 * 2 ** 20


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "sm_ghz": d.get("sm_max_mhz", 1965.0) / 1e3, "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_ghz": 1.965, "source": "fallback"}


class ClockSampler:
    """SM clock and throttle reasons sampled every 250 ms while the timed region runs.  NVML in a thread (a polling
    nvidia-smi process costs a launch-bound step several ms: its queries contend with the kernel launches);
    nvidia-smi is only the fallback when the NVML binding is missing."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self._stop = index, [], None, None, threading.Event()
        # 20 ms: a 20-step timed region lasts ~0.15 s and must still hold several samples (NVML reads are cheap)
        self.period = float(os.environ.get("BENCH_CLOCK_PERIOD", "0.02"))

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:  # noqa: BLE001
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "500"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def _poll(self):
        n = self.nvml
        bits = [(n.nvmlClocksThrottleReasonHwSlowdown, "hw_slowdown"),
                (n.nvmlClocksThrottleReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (n.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                (n.nvmlClocksThrottleReasonSwPowerCap, "sw_power_cap")]
        while not self._stop.is_set():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                flags = ["Active" if mask & b else "Not Active" for b, _ in bits]
                self.rows.append((time.time(), [str(sm), str(self.max_sm), "0"] + flags))
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(self.period)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t_begin=None, t_end=None):
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=2)
        elif self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ts, r in self.rows:
            if len(r) < 7 or (t_begin is not None and not (t_begin <= ts <= t_end + 0.15)):
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            for n, v in zip(self.NAMES, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def max_over_ranks(ms_local, dev, world):
    """Device-timed milliseconds -> max over ranks (the contract's multi-GPU timing rule)."""
    import torch.distributed as dist
    t = torch.tensor([ms_local], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(units_per_rank_step, steps, world, ms_total):
    """Whole-job units per second: every rank processed units_per_rank_step * steps in ms_total (max over ranks)."""
    return world * units_per_rank_step * steps / (ms_total * 1e-3)


def synthetic_images(batch, seed):
    from synthetic import synthetic_images as gen      # neutral module: neither the product nor the oracle
    return gen(batch, seed)


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_encoder_throughput(sample_images, steps, warmup):
    """images/s of the oracle's Encoder.forward restatement (fp32, all host threads)."""
    from oracle.encoder_oracle import encoder_forward, random_encoder_state
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.set_flush_denormal(True)   # exact-GELU tails make denormals: 14x slowdown otherwise
    sd = random_encoder_state(seed=0, layer_scale=1.0)
    x = synthetic_images(sample_images, 1234)
    with torch.no_grad():
        for _ in range(warmup):
            encoder_forward(sd, x, 7)
        t0 = time.perf_counter()
        for _ in range(steps):
            encoder_forward(sd, x, 7)
        dt = time.perf_counter() - t0
    return sample_images * steps / dt, cores, dt / steps


def reference_train_throughput(sample_images, steps, warmup, device="cpu", decoder="lstm", autocast=None,
                               finetune=True):
    """images/s of the reference's train step (trainMultiGPU.py:357-394) restated with stock torch ops — the oracle:
    torchvision-ConvNeXt arithmetic (frozen children 0-6 + trainable child 7, or fully frozen), LSTM-attention or
    Transformer decoder with teacher forcing, packed CE (+ alpha regulariser), backward, clamp +-5, Adam; train-mode
    dropout via injected masks.  device="cpu": the CPU baseline / --impl reference (fp32, all host threads);
    device=cuda: what PyTorch's own eager kernels (cuDNN / cuBLAS / ATen) make of the same step on this GPU."""
    from oracle import decoder_oracle as do
    from oracle import encoder_oracle as eo
    dev = torch.device(device)
    cores = os.cpu_count() or 1
    if dev.type == "cpu":
        torch.set_num_threads(cores)
        torch.set_flush_denormal(True)
    esd = eo.random_encoder_state(seed=0, layer_scale=1.0)
    dsd = do.random_lstm_decoder_state(0, V) if decoder == "lstm" else do.random_transformer_decoder_state(0, V)
    e_leaf = {k: v.to(dev).requires_grad_(finetune and k.startswith("convnext.7.")) for k, v in esd.items()}
    d_leaf = {k: v.to(dev).requires_grad_(v.is_floating_point() and k != "pos_encoding.pe") for k, v in dsd.items()}
    tr_e = [v for v in e_leaf.values() if v.requires_grad]
    tr_d = [v for v in d_leaf.values() if v.requires_grad]
    opt_e = torch.optim.Adam(tr_e, lr=1e-4) if tr_e else None
    opt_d = torch.optim.Adam(tr_d, lr=1e-4)
    imgs = synthetic_images(sample_images, 1234).to(dev)
    caps, lens = do.synthetic_captions(sample_images, 7, V)
    caps, lens = caps.to(dev), lens.to(dev)
    T = int(lens.max()) - 1

    def fwd_loss():
        if tr_e:
            feats = eo.encoder_forward(e_leaf, imgs, 7)
        else:
            with torch.no_grad():
                feats = eo.encoder_forward(e_leaf, imgs, 7)
        if decoder == "lstm":
            mask = (torch.rand(sample_images, T, 512, device=dev) > 0.5).float() * 2.0
            p, cs, dl, al, _ = do.lstm_teacher_forcing(d_leaf, feats, caps, lens, dropmask=mask)
            return do.train_loss_lstm(p, cs, dl, al)
        p, cs, dl = do.transformer_teacher_forcing(d_leaf, feats, caps, lens, caps == 0)
        return do.train_loss_transformer(p, cs, dl)

    def step():
        if autocast is not None:
            with torch.autocast(dev.type, dtype=autocast):
                loss = fwd_loss()
        else:
            loss = fwd_loss()
        if opt_e:
            opt_e.zero_grad()
        opt_d.zero_grad()
        loss.backward()
        for prm in tr_e + tr_d:
            if prm.grad is not None:
                prm.grad.clamp_(-5.0, 5.0)
        if opt_e:
            opt_e.step()
        opt_d.step()
        return loss

    for _ in range(warmup):
        step()
    if dev.type == "cuda":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = step()
    float(loss)
    if dev.type == "cuda":
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return sample_images * steps / dt, cores, dt / steps


def cpu_train_throughput(sample_images, steps, warmup):
    return reference_train_throughput(sample_images, steps, warmup, "cpu")


def cpu_single_image_caption(steps=2):
    """BASELINE.json configs[0]: caption.py-style single-image inference on CPU — Encoder.forward on one 256x256 image,
    then the beam search of caption.py:39-155 with beamSize = 1 (= greedy, caption.py:484) on the LSTM-attention
    decoder, random init, fp32, all host threads.  -> seconds per caption."""
    from oracle import decoder_oracle as do
    from oracle import encoder_oracle as eo
    torch.set_num_threads(os.cpu_count() or 1)
    torch.set_flush_denormal(True)
    esd = eo.random_encoder_state(seed=0, layer_scale=1.0)
    dsd = do.random_lstm_decoder_state(0, V, end_bias=3.2)
    img = synthetic_images(1, 77)
    ts, n_tok = [], 0
    with torch.no_grad():
        for i in range(steps + 1):
            t0 = time.perf_counter()
            feats = eo.encoder_forward(esd, img, 7)
            best, _, _ = do.beam_search(dsd, feats, "lstm", 1, V - 2, V - 1, V, max_steps=50)
            if i:
                ts.append(time.perf_counter() - t0)
                n_tok = len(best) if best is not None else 51
    return sum(ts) / len(ts), n_tok


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the same step on the same configuration (the full batch
    of 32 / 64 images per step), with all host threads, for exactly --steps steps after --warmup (capped at 2: a CPU
    step takes seconds) warm-up steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 30))      # a CPU step takes seconds: bounded so the run ends within minutes
    warmup = max(1, min(args.warmup, 2))
    if args.workload == "encoder":
        sample = args.batch
        ips, cores, spstep = cpu_encoder_throughput(sample, steps, warmup)
        metric, cfg = "encoder_forward_images_per_sec", workload_config(args)
        what = f"Encoder.forward on the full batch of {sample} synthetic 256x256 images per step"
    else:
        sample = TRAIN_BATCH
        ips, cores, spstep = cpu_train_throughput(sample, steps, warmup)
        metric, cfg = "train_images_per_sec", train_config(args, 1)
        what = f"one full train step (fwd, loss, bwd, clamp, Adam) on the full batch of {sample} synthetic images per step"
    line = {
        "impl": "reference", "metric": metric, "value": ips, "unit": "images/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": spstep * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{what}, {steps} steps, fp32, torch {torch.__version__} CPU ops, "
                                   f"flush-denormal on"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def train_config(args, world, sample_note=None):
    c = {"workload": "trainMultiGPU.py train step (BASELINE.json configs[3]): Encoder.fine_tune(True, startingLayer=7) "
                     "+ DecoderWithAttention, teacher forcing, batch 32 per GPU, 256x256 synthetic images, 52-token "
                     "caption rows (lengths uniform 7..52), packed CE + alpha regulariser, clamp +-5, Adam",
         "batch_per_gpu": TRAIN_BATCH, "image": "3x256x256", "caption_tokens": 52, "vocab": V,
         "weights": "random init (reference initialisers, seed 0), layer_scale=1.0; dropout 0.5 and stochastic depth on",
         "parallelism": f"dp{world}" + (" (DistributedDataParallel, NCCL all-reduce overlapped with backward)"
                                        if world > 1 else ""),
         "l2_policy": "4 input batches rotated; one step touches > 1 GB of activations / weights / Adam state "
                      "(L2 = 126 MB), so nothing survives in L2 between steps"}
    if sample_note:
        c["note"] = sample_note
    return c


def workload_config(args, sample_note=None):
    c = {"workload": f"Encoder.forward (ConvNeXt-Base features + AdaptiveAvgPool 7x7), batch {args.batch} per GPU, "
                     f"256x256 synthetic images (BASELINE.json configs[1])",
         "batch_per_gpu": args.batch, "image": "3x256x256", "encoded_image_size": 7,
         "weights": "random init (torchvision initialiser, seed 0), layer_scale=1.0",
         "l2_policy": "4 input batches rotated (201 MB > 126 MB L2); activations per step (>2 GB) exceed L2",
         "launch": "eager" if getattr(args, "no_graph", False) else "CUDA-graph replay of the forward (Encoder.enable_cuda_graph)"}
    if sample_note:
        c["note"] = sample_note
    return c


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Ctx:
    """Process-group / device context of one bench process."""

    def __init__(self, args):
        import torch.distributed as dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torchrun (one rank per GPU)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            opts = None
            if os.environ.get("CCX_NCCL_HIGH_PRIO", "0") != "0":
                # the gradient all-reduce runs next to the rest of the backward (CapturedTrainStep): on a high-priority
                # stream its CTAs are placed ahead of the already queued CTAs of the compute kernels
                opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            dist.init_process_group("nccl", device_id=self.dev, pg_options=opts)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def close(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


def run_ours(args):
    ctx = Ctx(args)
    if args.workload == "encoder":
        line = encoder_line(args, ctx, with_cpu=not args.no_cpu_baseline)
        if line is not None and not args.no_extras:
            line["extra"] = _guard(lambda: run_extras(args, ctx))
    else:
        line = train_line(args, ctx)
        extra = None
        if not args.no_extras:
            torch.cuda.empty_cache()
            sub = argparse.Namespace(**vars(args))
            sub.steps, sub.warmup, sub.spans = 50, 5, None
            extra = {}
            enc_line = _guard(lambda: encoder_line(sub, ctx, with_cpu=False))
            if ctx.rank == 0:
                keep = ("metric", "value", "unit", "ms_per_step", "dtype", "config", "e2e", "gpu_launches_per_step",
                        "roofline", "kernels", "model_tflops", "error")
                extra["encoder_forward_configs1"] = {k: enc_line[k] for k in keep if k in enc_line}
            torch.cuda.empty_cache()
            args._headline_lstm_ips = (line["value"] / ctx.world) if line is not None else None
            more = _guard(lambda: run_extras(args, ctx))
            extra.update(more or {})
        if line is not None:
            line["extra"] = extra
            te = (extra or {}).get("torch_eager_b200") or {}
            if "train_lstm_finetune7" in te and "x_over_torch_eager" in te["train_lstm_finetune7"]:
                line["x_over_torch_eager"] = te["train_lstm_finetune7"]["x_over_torch_eager"]
    if ctx.rank == 0:
        emit(line)
    ctx.close()


def _guard(fn):
    """The headline line must survive a failing secondary workload."""
    try:
        return fn()
    except Exception as ex:  # noqa: BLE001
        return {"error": f"{type(ex).__name__}: {ex}"}


def encoder_line(args, ctx, with_cpu):
    """BASELINE.json configs[1]: Encoder.forward, batch 64 per GPU -> the JSON line (rank 0) / None (other ranks)."""
    import torch.distributed as dist
    from imagecaptioningconvnext_b200 import Encoder, _lib
    from synthetic import random_encoder_state
    world, rank, local, dev, barrier = ctx.world, ctx.rank, ctx.local, ctx.dev, ctx.barrier

    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    enc = Encoder(encoded_image_size=7, compute_dtype=dtype)
    enc.load_state_dict(random_encoder_state(seed=0, layer_scale=1.0))
    enc = enc.to(dev).eval()
    if not args.no_graph:
        enc.enable_cuda_graph()      # public inference option: the 117-launch forward replayed as one CUDA graph

    B = args.batch
    nbuf = 4
    host = [synthetic_images(B, 1234 + rank * 16 + i).pin_memory() for i in range(nbuf)]
    devbuf = [h.to(dev) for h in host]

    # ---- device-resident throughput -------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # started early: nvidia-smi needs ~1 s before its first sample
    with torch.no_grad():
        for i in range(args.warmup):
            enc(devbuf[i % nbuf])
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_begin = time.time()
        e0.record()
        for i in range(args.steps):
            out = enc(devbuf[i % nbuf])
        e1.record()
        barrier()
        t_end = time.time()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    ms_total = max_over_ranks(ms, dev, world)
    value = aggregate_throughput(B, args.steps, world, ms_total)

    # ---- end to end through the public call, host buffers ----------------------------------------
    # Every step uploads ITS images from pinned host memory and downloads ITS features inside the timed region.
    # (a) serial: copy -> forward -> copy on one stream;  (b) pipelined: the next batch's upload and the previous
    # batch's download run on a copy stream while the current batch computes (what an inference loop would do).
    with torch.no_grad():
        stage = torch.empty_like(devbuf[0])
        res_host = torch.empty((B, 7, 7, 1024), dtype=torch.float32).pin_memory()
        for i in range(2):
            stage.copy_(host[i % nbuf], non_blocking=True)
            res_host.copy_(enc(stage), non_blocking=True)
        barrier()
        e0.record()
        for i in range(args.steps):
            stage.copy_(host[i % nbuf], non_blocking=True)        # H2D of this step's images (pinned)
            res_host.copy_(enc(stage), non_blocking=True)         # D2H of this step's features
        e1.record()
        barrier()
        ms_e2e_serial = e0.elapsed_time(e1)

        main = torch.cuda.current_stream()
        copy_s = torch.cuda.Stream()
        stages = [torch.empty_like(devbuf[0]) for _ in range(2)]
        results = [torch.empty((B, 7, 7, 1024), dtype=torch.float32).pin_memory() for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        barrier()
        e0.record()
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(e0)
            stages[0].copy_(host[0], non_blocking=True)
            ready[0].record(copy_s)
        for i in range(args.steps):
            cur, nxt = i % 2, (i + 1) % 2
            if i + 1 < args.steps:
                with torch.cuda.stream(copy_s):
                    if i >= 1:
                        copy_s.wait_event(consumed[nxt])          # the forward that read this stage has finished
                    stages[nxt].copy_(host[(i + 1) % nbuf], non_blocking=True)
                    ready[nxt].record(copy_s)
            main.wait_event(ready[cur])
            out = enc(stages[cur])
            consumed[cur].record(main)
            with torch.cuda.stream(copy_s):
                copy_s.wait_event(consumed[cur])
                results[cur].copy_(out, non_blocking=True)
                out.record_stream(copy_s)
        main.wait_stream(copy_s)
        e1.record()
        barrier()
        ms_e2e = e0.elapsed_time(e1)
    e2e_value = aggregate_throughput(B, args.steps, world, max_over_ranks(ms_e2e, dev, world))
    e2e_serial = aggregate_throughput(B, args.steps, world, max_over_ranks(ms_e2e_serial, dev, world))

    # ---- instrumented pass: per-kernel CUDA events (same steps, same data) -------------------------
    enc.enable_cuda_graph(False)     # per-launch events need the eager launch path
    with torch.no_grad():
        torch.cuda.synchronize()
        _lib.prof_begin()
        for i in range(args.steps):
            enc(devbuf[i % nbuf])
        spans = _lib.prof_spans() if args.spans else None
        prof = _lib.prof_end()
    if spans is not None and rank == 0:
        per = len(spans) // args.steps
        with open(args.spans, "w") as f:
            f.write("# launch index within one step, kind, mean ms over steps, work (FLOPs for gemm, bytes otherwise), rate\n")
            for i in range(per):
                ms_i = sum(spans[s * per + i][1] for s in range(args.steps)) / args.steps
                k, _, wk = spans[i]
                rate = wk / (ms_i * 1e-3) if ms_i > 0 else 0
                f.write(f"{i:4d} {k:12s} {ms_i * 1e3:9.1f} us  work={wk:.4g}  "
                        f"{rate / 1e12:8.1f} {'TFLOP/s' if k == 'gemm' else 'TB/s'}\n")
    if world > 1:
        dist.barrier()

    del enc
    if rank != 0:
        return None

    peaks = measured_peaks()
    tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
    kernels = {}
    for k, v in prof.items():
        if v["launches"] == 0:
            continue
        kernels[k] = kernel_row(k, v, args.steps, tot_ms, peaks)
    g = prof["gemm"]
    achieved = g["work"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
    peak = peaks["bf16_tflops_sustained"] if dtype == torch.bfloat16 else peaks["bf16_tflops_sustained"] / 6.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath) and dtype == torch.bfloat16:
        traffic = json.load(open(tpath))["gemm_tn_kernel"]["dram_bytes_per_launch"]   # from the committed ncu capture
    roofline = {"kernel": "gemm_tn_kernel (tcgen05, all encoder pointwise/downsample GEMMs)", "bound": "tensor",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peaks["source"] + (" bf16 sustained" if dtype == torch.bfloat16
                                                                     else " bf16 sustained / 6 (3xTF32 at half rate)"),
                "flops_per_launch": g["work"] / max(g["launches"], 1),
                "us_per_launch": g["ms"] * 1e3 / max(g["launches"], 1)}
    launches = sum(v["launches"] for v in prof.values()) // args.steps

    cpu = None
    if world == 1 and with_cpu:
        sample, csteps = 4, 3
        ips, cores, _ = cpu_encoder_throughput(sample, csteps, 1)
        cpu = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"oracle Encoder.forward on {sample} of the {B} images x {csteps} steps, fp32, all host threads"}

    line = {
        "metric": "encoder_forward_images_per_sec", "value": value, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if dtype == torch.bfloat16 else "fp32(3xTF32)", "data": "synthetic",
        "config": workload_config(args),
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": B * 3 * 256 * 256 * 4,
                "d2h_bytes_per_step": B * 49 * 1024 * 4, "serial_value": e2e_serial,
                "how": "Encoder.__call__ per step on host-pinned fp32 images; H2D of batch i+1 and D2H of batch i-1 "
                       "overlap the forward of batch i on a copy stream (serial_value: everything on one stream)"},
        "gpu_launches": int(launches * args.steps),
        "gpu_launches_per_step": int(launches),
        "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu, "clocks": clocks,
        "model_tflops": value / world * ENCODER_GFLOP_PER_IMAGE / 1e3,
    }
    return line


def comm_attribution(ctx, enc, dec, cap_step, devbuf, ms_step):
    """N > 1: where does the step time beyond the single-GPU step go?  (a) the gradient all-reduces of one step timed
    alone, back to back, on an idle GPU; (b) the same captured step with the all-reduces left out.  exposed = step -
    step without all-reduce; hidden fraction = 1 - exposed / (a)."""
    import torch.distributed as dist
    from imagecaptioningconvnext_b200.train_step import CapturedTrainStep, make_optimizers
    dev, world = ctx.dev, ctx.world
    slices = [cap_step._buckets[0]]
    if len(cap_step._buckets) > 1:
        slices += [cap_step._buckets[1][a:b] for a, b in cap_step._enc_units.values()]

    def all_reduces():
        for sl in slices:
            dist.all_reduce(sl, op=dist.ReduceOp.AVG)
    ms_ar = _timed(all_reduces, 10, 3, dev, world)
    d_opt2, e_opt2 = make_optimizers(enc, dec)
    quiet = CapturedTrainStep(enc, dec, d_opt2, e_opt2, skip_allreduce=True)
    ms_quiet = _timed(lambda: quiet(*devbuf[0]), 20, 8, dev, world)
    exposed = max(ms_step - ms_quiet, 0.0)
    return {"bytes_per_step": int(sum(sl.numel() for sl in slices) * 4), "messages_per_step": len(slices),
            "ms_alone_back_to_back": ms_ar, "ms_step_without_allreduce": ms_quiet, "ms_step": ms_step,
            "exposed_ms": exposed, "hidden_fraction": 1.0 - min(exposed / ms_ar, 1.0) if ms_ar > 0 else None,
            "algorithm_bandwidth_gbs": sum(sl.numel() for sl in slices) * 4 / (ms_ar * 1e-3) / 1e9,
            "how": "fp32 buckets, NCCL average all-reduce inside the step's CUDA graph: decoder bucket after the "
                   "decoder backward, one slice per CNBlock of the fine-tuned stage as each block's backward ends"}


def count_kernels(step_fn, n):
    """GPU kernels per call of step_fn(i), counted by CUPTI (torch.profiler) — library, ATen and NCCL kernels alike,
    whether launched eagerly or from a CUDA-graph replay."""
    from torch.profiler import ProfilerActivity, profile
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(n):
            step_fn(i)
        torch.cuda.synchronize()
    kinds = ("kernel",)
    cnt = sum(1 for e in prof.events() if str(getattr(e, "device_type", "")).endswith("CUDA")
              and not e.name.lower().startswith(("memcpy", "memset")))
    return cnt / n


# what bounds each kernel kind of the library's per-launch timing (ccx_prof): tensor pipe or HBM
_TENSOR_KINDS = ("gemm", "gemm_skinny")
_LATENCY_KINDS = {"lstm": "one persistent cooperative kernel per direction: 51 dependent steps x 3 hand-overs, "
                          "latency-bound (no roofline claim); work column = recurrent GEMM FLOPs",
                  "gemm_skinny": "M <= 32 rows: latency-bound", "attention": "per-sample, latency-bound"}


def kernel_row(kind, v, n_steps, tot_ms, peaks):
    rate = v["work"] / (v["ms"] * 1e-3) if v["ms"] > 0 else 0.0
    row = {"launches_per_step": v["launches"] / n_steps, "ms_per_step": v["ms"] / n_steps, "share": v["ms"] / tot_ms}
    if kind in _TENSOR_KINDS or kind == "lstm":
        row["tflops"] = rate / 1e12
        row["frac_of_tensor_peak"] = rate / 1e12 / peaks["bf16_tflops_sustained"]
    else:
        row["gbs"] = rate / 1e9
        row["frac_of_hbm_peak"] = rate / 1e9 / peaks["hbm_gbs"]
    if kind == "dwconv_ln":
        # 49 MACs per output element: the FP32 FMA pipe, not HBM, is this kernel's roofline (per element the pipe needs
        # 49 / (148 SMs x 128 lanes) cycles = 1.36 ps at 1.9 GHz against 6 bytes / 6.5 TB/s = 0.92 ps of HBM time);
        # work column = bytes at 6 per element (fp32 in, bf16 out), so MACs = bytes / 6 x 49 (a lower bound for the
        # fp32-output backward calls)
        fp32_peak = 148 * 128 * 2 * peaks.get("sm_ghz", 1.9) * 1e9
        row["fp32_tflops"] = rate / 6.0 * 49 * 2 / 1e12
        row["frac_of_fp32_pipe_peak"] = rate / 6.0 * 49 * 2 / fp32_peak
        row["note"] = "bound by the FP32 FMA pipe (49 MACs per output), see frac_of_fp32_pipe_peak; frac_of_hbm_peak kept for reference"
    if kind in _LATENCY_KINDS:
        row["note"] = _LATENCY_KINDS[kind]
    return row


def train_line(args, ctx):
    """BASELINE.json configs[3] — the trainMultiGPU.py step — as the headline JSON line (rank 0) / None."""
    from torch.nn.parallel import DistributedDataParallel as DDP
    from imagecaptioningconvnext_b200 import DecoderWithAttention, Encoder, _lib
    from imagecaptioningconvnext_b200.train_step import CapturedTrainStep, caption_train_step, make_optimizers
    from synthetic import random_encoder_state, random_lstm_decoder_state, synthetic_captions
    world, rank, local, dev, barrier = ctx.world, ctx.rank, ctx.local, ctx.dev, ctx.barrier
    B, bf16 = TRAIN_BATCH, torch.bfloat16
    torch.manual_seed(42 + rank)                      # trainMultiGPU.py:8: per-rank dropout / stochastic-depth streams
    enc = Encoder(compute_dtype=bf16)
    enc.load_state_dict(random_encoder_state(seed=0, layer_scale=1.0))
    enc = enc.to(dev).train()
    enc.fine_tune(True, 7)                            # trainMultiGPU.py:68,223: startingLayer=7
    dec = DecoderWithAttention(512, 512, 512, V, dev, compute_dtype=bf16)
    dec.load_state_dict(random_lstm_decoder_state(0, V))
    dec = dec.to(dev).train()
    d_opt, e_opt = make_optimizers(enc, dec)
    captured = not args.eager_step
    if captured:
        # the repo's public train-step call: the whole step replayed as ONE CUDA graph; for N > 1 its two flat
        # gradient buckets are all-reduced over NCCL inside the graph (decoder bucket while the encoder stage still
        # back-propagates) — this replaces DistributedDataParallel's reducer (trainMultiGPU.py:233-236)
        enc_w, dec_w = enc, dec
        cap_step = CapturedTrainStep(enc, dec, d_opt, e_opt)
    else:
        ddp_kw = json.loads(os.environ.get("BENCH_DDP_KW", "{}"))     # experiments only; default = the reference's call
        enc_w = DDP(enc, device_ids=[local], **ddp_kw) if world > 1 else enc      # trainMultiGPU.py:233-236
        dec_w = DDP(dec, device_ids=[local], **ddp_kw) if world > 1 else dec

    nbuf = 4
    host = []
    for i in range(nbuf):
        caps, lens = synthetic_captions(B, 7 + rank * 16 + i, V)
        host.append((synthetic_images(B, 1234 + rank * 16 + i).pin_memory(), caps.pin_memory(), lens.pin_memory()))
    devbuf = [tuple(t.to(dev) for t in h) for h in host]

    def step(batch, host_lens=None):
        # host_lens: the lengths as the data loader yielded them (host memory); the device copy is batch[2]
        if captured:
            return cap_step(batch[0], batch[1], batch[2])
        return caption_train_step(enc_w, dec_w, batch[0], batch[1], batch[2], d_opt, e_opt, caplens_host=host_lens)

    # set-up (not warm-up): lazily built caches, the allocator's pools and — for the captured step — its eager
    # rehearsal steps and the graph capture itself; the --warmup steps after it run exactly like the timed ones
    for i in range(6):
        step(devbuf[i % nbuf], host[i % nbuf][2])
    warmup = args.warmup
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- device-resident throughput -------------------------------------------------------------
    for i in range(warmup):
        step(devbuf[i % nbuf], host[i % nbuf][2])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.time()
    e0.record()
    for i in range(args.steps):
        loss = step(devbuf[i % nbuf], host[i % nbuf][2])
    e1.record()
    host_ms = (time.time() - t_begin) * 1e3 / args.steps     # host time to ENQUEUE a step (no sync inside the loop)
    barrier()
    t_end = time.time()
    ms_total = max_over_ranks(e0.elapsed_time(e1), dev, world)
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    value = aggregate_throughput(B, args.steps, world, ms_total)
    last_loss = float(loss)

    # ---- end to end: host-pinned batches in, loss out, every step ---------------------------------
    # the next batch is uploaded on a copy stream while the current step computes (a DataLoader with pin_memory and
    # non_blocking copies, train.py:152-155,257-259, does the same); the loss is read back to pinned memory per step
    main, copy_s = torch.cuda.current_stream(), torch.cuda.Stream()
    stages = [tuple(torch.empty_like(t) for t in devbuf[0]) for _ in range(2)]
    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def upload(slot, i):
        for d, h in zip(stages[slot], host[i % nbuf]):
            d.copy_(h, non_blocking=True)

    barrier()
    e0.record()
    with torch.cuda.stream(copy_s):
        copy_s.wait_event(e0)
        upload(0, 0)
        ready[0].record(copy_s)
    for i in range(args.steps):
        cur, nxt = i % 2, (i + 1) % 2
        if i + 1 < args.steps:
            with torch.cuda.stream(copy_s):
                if i >= 1:
                    copy_s.wait_event(consumed[nxt])
                upload(nxt, i + 1)
                ready[nxt].record(copy_s)
        main.wait_event(ready[cur])
        loss = step(stages[cur], host[i % nbuf][2])
        consumed[cur].record(main)
        loss_host[cur:cur + 1].copy_(loss.reshape(1), non_blocking=True)        # D2H of this step's loss
    e1.record()
    barrier()
    e2e_value = aggregate_throughput(B, args.steps, world, max_over_ranks(e0.elapsed_time(e1), dev, world))
    h2d = sum(t.numel() * t.element_size() for t in host[0])

    # ---- instrumented pass: per-kernel CUDA events over the same steps -----------------------------
    # true number of GPU kernels of one step (library kernels + the ATen index/fill/RNG kernels around them + NCCL),
    # counted by CUPTI through torch.profiler around two steps of the SAME call that was timed
    kernels_per_step = count_kernels(lambda i: step(devbuf[i % nbuf], host[i % nbuf][2]), 2)

    def eager(batch):     # per-launch events need the eager launch path (a graph replay bypasses the library hooks)
        if captured:
            return cap_step.eager_step(batch[0], batch[1], batch[2])
        return step(batch, None)
    n_inst = min(args.steps, 10)
    torch.cuda.synchronize()
    _lib.prof_begin()
    for i in range(n_inst):
        eager(devbuf[i % nbuf])
    spans = _lib.prof_spans() if args.spans else None
    prof = _lib.prof_end()
    if spans is not None and rank == 0:
        per = len(spans) // n_inst
        with open(args.spans, "w") as f:
            f.write("# launch index within the first instrumented step, kind, ms, work (FLOPs for gemm, bytes otherwise)\n")
            for i, (k, ms_i, wk) in enumerate(spans[:per]):
                f.write(f"{i:4d} {k:12s} {ms_i * 1e3:9.1f} us  work={wk:.4g}\n")
    barrier()
    comm = None
    if world > 1 and captured:
        comm = comm_attribution(ctx, enc, dec, cap_step, devbuf, ms_total / args.steps)
    del enc_w, dec_w, d_opt, e_opt, enc, dec, devbuf, stages
    if captured:
        del cap_step
    if rank != 0:
        return None

    peaks = measured_peaks()
    tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
    kernels = {}
    for k, v in prof.items():
        if v["launches"] == 0:
            continue
        kernels[k] = kernel_row(k, v, n_inst, tot_ms, peaks)
    g = prof["gemm"]
    achieved = g["work"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
    peak = peaks["bf16_tflops_sustained"]
    sk = prof.get("gemm_skinny", {"ms": 0.0, "work": 0.0, "launches": 0})
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):      # committed ncu --set full capture of this kernel inside this bench command
        traffic = json.load(open(tpath)).get("gemm_tn_kernel_train", {}).get("dram_bytes_per_launch")
    roofline = {"kernel": "gemm_tn_kernel (tcgen05/TMEM, TMA-fed): the M >= 33 GEMMs of the step — encoder pointwise / "
                          "downsample forward at B=32, stage-4 dgrad / wgrad, hoisted attention and vocabulary "
                          "projections, time-batched weight gradients",
                "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peaks["source"] + " bf16 sustained",
                "flops_per_launch": g["work"] / max(g["launches"], 1),
                "us_per_launch": g["ms"] * 1e3 / max(g["launches"], 1),
                "share_of_step_kernel_time": g["ms"] / tot_ms,
                "note": "largest kernel by time share; traffic = DRAM bytes per launch of two sampled stage-3 launches "
                        "(profiles/ncu_traffic.json; algorithmic bytes of the same launches: 56.6 MB).  The M <= 32 recurrent GEMMs run on gemm_skinny_kernel "
                        "(mma.sync, latency-bound, no roofline claim): "
                        f"{sk['launches'] / n_inst:.0f} launches/step, {sk['ms'] / n_inst:.2f} ms/step, "
                        f"{sk['ms'] * 1e3 / max(sk['launches'], 1):.1f} us each"}
    launches = sum(v["launches"] for v in prof.values()) // n_inst
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        csteps = 3
        ips, cores, _ = cpu_train_throughput(B, csteps, 1)
        cpu = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"oracle train step (fwd, loss, bwd, clamp, Adam) on the same full batch of {B} images, "
                         f"{csteps} steps after 1 warm-up, fp32, all host threads (the same run --impl reference times)"}
    return {
        "metric": "train_images_per_sec", "value": value, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": train_config(args, world),
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "how": ("CapturedTrainStep" if captured else "caption_train_step") + " per step on host-pinned images "
                       "/ captions / lengths (upload of batch i+1 on a copy stream during step i) and a D2H read of "
                       "every step's loss"},
        "gpu_launches": int(kernels_per_step * args.steps), "gpu_launches_per_step": int(kernels_per_step),
        "library_launches_per_step": int(launches),
        "step_call": "CapturedTrainStep (one CUDA-graph replay per step)" if captured else "caption_train_step (eager)",
        "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu, "clocks": clocks,
        "last_loss": last_loss,
        "host_enqueue_ms_per_step": host_ms,
        "allreduce": comm,
    }


# ------------------------------------------------------------------------------------------------
# secondary workloads (BASELINE.json configs[2..4]) reported under "extra" in the same JSON line
# ------------------------------------------------------------------------------------------------
def _timed(fn, steps, warmup, dev, world):
    import torch.distributed as dist
    for _ in range(warmup):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / steps


def run_extras(args, ctx):
    """Transformer train images/s (configs[2]) and beam-search captions/s (configs[4]); device-resident inputs.
    With --workload encoder also the LSTM train step (configs[3], otherwise the headline)."""
    dev, world, rank, local = ctx.dev, ctx.world, ctx.rank, ctx.local
    from torch.nn.parallel import DistributedDataParallel as DDP
    from imagecaptioningconvnext_b200 import DecoderWithAttention, Encoder, TransformerDecoder
    from imagecaptioningconvnext_b200.beam import CapturedBeamSearch
    from imagecaptioningconvnext_b200.train_step import caption_train_step, make_optimizers
    from synthetic import (random_encoder_state, random_lstm_decoder_state, random_transformer_decoder_state,
                           synthetic_captions)
    out = {}
    B = 32
    bf16 = torch.bfloat16
    esd = random_encoder_state(seed=0, layer_scale=1.0)
    imgs = synthetic_images(B, 99 + rank).to(dev)
    caps, lens = synthetic_captions(B, 7 + rank, V)
    caps, lens = caps.to(dev), lens.to(dev)

    def wrap(m):
        return DDP(m, device_ids=[local]) if world > 1 and any(p.requires_grad for p in m.parameters()) else m

    # configs[3]: encoder fine-tuned from child 7 + LSTM-attention decoder, bf16, DDP (the headline unless --workload encoder)
    if args.workload == "encoder":
        enc = Encoder(compute_dtype=bf16)
        enc.load_state_dict(esd)
        enc = enc.to(dev).train()
        enc.fine_tune(True, 7)
        dec = DecoderWithAttention(512, 512, 512, V, dev, compute_dtype=bf16)
        dec.load_state_dict(random_lstm_decoder_state(0, V))
        dec = dec.to(dev).train()
        d_opt, e_opt = make_optimizers(enc, dec)
        enc_w, dec_w = wrap(enc), wrap(dec)
        ms = _timed(lambda: caption_train_step(enc_w, dec_w, imgs, caps, lens, d_opt, e_opt), 20, 10, dev, world)
        out["train_lstm_finetune7_bf16"] = {"images_per_sec": world * B / (ms * 1e-3), "ms_per_step": ms,
                                            "batch_per_gpu": B, "config": "BASELINE.json configs[3]: encoder "
                                            "fine_tune(True,7) + DecoderWithAttention, teacher forcing, captions uniform "
                                            "7..52 tokens, dropout/stochastic depth on, clamp+Adam"
                                            + (", DDP/NCCL" if world > 1 else "")}
        del enc_w, dec_w, d_opt, e_opt, dec
    # configs[2]: frozen encoder + TransformerDecoder teacher forcing
    for name, cd in (("bf16", bf16), ("fp32", torch.float32)):
        enc2 = Encoder(compute_dtype=cd)
        enc2.load_state_dict(esd)
        enc2 = enc2.to(dev).train()
        enc2.fine_tune(False)
        tr = TransformerDecoder(512, 512, V, 52, dev, None, None, True, compute_dtype=cd)
        tr.load_state_dict(random_transformer_decoder_state(0, V))
        tr = tr.to(dev).train()
        d_opt, _ = make_optimizers(enc2, tr)
        tr_w = wrap(tr)
        ms_eager = _timed(lambda: caption_train_step(enc2, tr_w, imgs, caps, lens, d_opt, None), 20, 10, dev, world)
        tr.enable_cuda_graph()       # decoder forward / backward bodies replayed as two CUDA graphs
        ms_two_graphs = _timed(lambda: caption_train_step(enc2, tr_w, imgs, caps, lens, d_opt, None), 20, 10, dev, world)
        tr.enable_cuda_graph(False)
        del tr_w, d_opt
        # the whole step (frozen encoder forward, decoder forward / loss / backward, all-reduce, clamp+Adam) as ONE graph
        from imagecaptioningconvnext_b200.train_step import CapturedTrainStep
        d_opt, _ = make_optimizers(enc2, tr)
        cap = CapturedTrainStep(enc2, tr, d_opt, None)
        ms = _timed(lambda: cap(imgs, caps, lens), 20, 10, dev, world)
        out[f"train_transformer_frozen_encoder_{name}"] = {
            "images_per_sec": world * B / (ms * 1e-3), "ms_per_step": ms, "ms_per_step_eager_launches": ms_eager,
            "ms_per_step_decoder_graphs_only": ms_two_graphs, "batch_per_gpu": B,
            "step_call": "CapturedTrainStep (one CUDA-graph replay per step)",
            "config": "BASELINE.json configs[2]: frozen encoder + TransformerDecoder, teacher forcing, 52-token "
                      "rows (captions uniform 7..52), dropout on, clamp+Adam" + (", NCCL all-reduce in the graph" if world > 1 else "")}
        del cap, d_opt
    # trainWithoutTeacherForcing (trainMultiGPU.py:423-498): greedy generation + one
    # differentiable pass over the generated ids; random-init decoders never emit <end>, i.e. all 51 steps run
    for kind in ("lstm", "transformer"):
        encf = Encoder(compute_dtype=bf16)
        encf.load_state_dict(esd)
        encf = encf.to(dev).train()
        encf.fine_tune(False)
        if kind == "lstm":
            decf = DecoderWithAttention(512, 512, 512, V, dev, compute_dtype=bf16)
            decf.load_state_dict(random_lstm_decoder_state(0, V))
        else:
            decf = TransformerDecoder(512, 512, V, 52, dev, None, None, True, compute_dtype=bf16)
            decf.load_state_dict(random_transformer_decoder_state(0, V))
        decf = decf.to(dev).train()
        d_opt, _ = make_optimizers(encf, decf)
        dec_w = wrap(decf)
        ms = _timed(lambda: caption_train_step(encf, dec_w, imgs, caps, lens, d_opt, None, teacher_forcing=False,
                                               wordMap=WORDMAP), 10, 5, dev, world)
        out[f"train_free_running_{kind}_frozen_encoder_bf16"] = {
            "images_per_sec": world * B / (ms * 1e-3), "ms_per_step": ms, "batch_per_gpu": B,
            "config": "trainWithoutTeacherForcing step: frozen encoder, 51 free-running steps (KV cache for the "
                      "Transformer), loss on the generated positions, clamp+Adam" + (", DDP/NCCL" if world > 1 else "")}
        del dec_w, d_opt, decf, encf
    # configs[4]: batched beam search k=5, 128 images per GPU, TransformerDecoder (encoder included)
    NI = 128
    enc3 = Encoder(compute_dtype=bf16)
    enc3.load_state_dict(esd)
    enc3 = enc3.to(dev).eval()
    tr = TransformerDecoder(512, 512, V, 52, dev, None, None, True, compute_dtype=bf16)
    tr.load_state_dict(random_transformer_decoder_state(0, V, end_bias=3.2))
    tr = tr.to(dev).eval()
    big = synthetic_images(NI, 5 + rank).to(dev)

    searcher = CapturedBeamSearch(tr, WORDMAP, "transformer", beamSize=5)

    def beam():
        with torch.no_grad():
            return searcher(enc3(big))          # encoder + CUDA-graph replay of the 51-step decode + D2H of results
    ms = _timed(beam, 4, 2, dev, world)
    out["beam_search_transformer_k5_bf16"] = {"captions_per_sec": world * NI / (ms * 1e-3), "ms_per_batch": ms,
                                              "images_per_gpu": NI, "beam": 5, "max_steps": 51,
                                              "config": "BASELINE.json configs[4]: Encoder + TransformerDecoder beam "
                                                        "search (KV cache, decode loop replayed as one CUDA graph), "
                                                        "captions read back to the host, replicas only, no collective"}
    # "reported beside torch-eager-on-B200" — the stock torchvision ConvNeXt-base feature stack (what
    # models/encoder.py:18-20 wraps) + AdaptiveAvgPool2d + permute, run by PyTorch's own kernels on this GPU.  Library
    # comparison only; rank 0, same B=64 batch shape as the headline.
    if rank == 0:
        try:
            out["torch_eager_encoder_b200"] = _torch_eager_encoder(dev)
        except Exception as ex:  # noqa: BLE001  (torchvision missing / OOM: report, do not fail the bench)
            out["torch_eager_encoder_b200"] = {"error": f"{type(ex).__name__}: {ex}"}
        del enc3, tr, searcher, big
        torch.cuda.empty_cache()
        ours = {"transformer": out["train_transformer_frozen_encoder_bf16"]["images_per_sec"] / world,
                "transformer_fp32": out["train_transformer_frozen_encoder_fp32"]["images_per_sec"] / world,
                "beam": out["beam_search_transformer_k5_bf16"]["captions_per_sec"] / world}
        if "train_lstm_finetune7_bf16" in out:
            ours["lstm"] = out["train_lstm_finetune7_bf16"]["images_per_sec"] / world
        elif getattr(args, "_headline_lstm_ips", None):
            ours["lstm"] = args._headline_lstm_ips
        out["torch_eager_b200"] = _guard(lambda: _torch_eager_arms(dev, ours))
        if world == 1 and not args.no_cpu_baseline:
            def c1():
                sec, ntok = cpu_single_image_caption(2)
                return {"seconds_per_caption": sec, "captions_per_sec": 1.0 / sec, "tokens": ntok,
                        "cores": os.cpu_count(), "kind": "port",
                        "config": "BASELINE.json configs[0]: caption.py-style single-image inference on CPU — ConvNeXt "
                                  "encoder + LSTM-attention decoder, random init, 256x256 synthetic image, beamSize=1 "
                                  "(greedy, caption.py:484), fp32, all host threads"}
            out["cpu_single_image_caption_configs0"] = _guard(c1)
    return out


def _torch_eager_arms(dev, ours):
    """"the bar to beat is PyTorch-eager on the same B200".  The reference's own step bodies
    (the oracle's stock-torch restatement of trainMultiGPU.py:357-394 and caption.py:160-255, which is what the
    reference modules execute) run by PyTorch's eager CUDA kernels — cuDNN / cuBLAS / ATen — on this GPU, same batch,
    same shapes, fp32 (the reference's only precision; TF32 left at torch's defaults) and bf16 autocast.  Per GPU."""
    res = {"what": "reference step bodies on torch " + torch.__version__ + " eager CUDA kernels, same B200, batch 32"}

    def train(decoder, finetune):
        r = {}
        for name, ac in (("fp32", None), ("bf16_autocast", torch.bfloat16)):
            ips, _, sps = reference_train_throughput(TRAIN_BATCH, 5, 2, dev, decoder, ac, finetune)
            r[name] = {"images_per_sec": ips, "ms_per_step": sps * 1e3}
            torch.cuda.empty_cache()
        r["best_images_per_sec"] = max(v["images_per_sec"] for v in r.values())
        return r

    res["train_lstm_finetune7"] = train("lstm", True)
    res["train_transformer_frozen_encoder"] = train("transformer", False)
    if "lstm" in ours:
        res["train_lstm_finetune7"]["x_over_torch_eager"] = ours["lstm"] / res["train_lstm_finetune7"]["best_images_per_sec"]
    res["train_transformer_frozen_encoder"]["x_over_torch_eager"] = \
        ours["transformer"] / res["train_transformer_frozen_encoder"]["best_images_per_sec"]
    res["train_transformer_frozen_encoder"]["x_over_torch_eager_fp32_vs_fp32"] = \
        ours["transformer_fp32"] / res["train_transformer_frozen_encoder"]["fp32"]["images_per_sec"]
    # beam search k=5, Transformer decoder, image by image as caption.py does (whole-prefix recompute, no KV cache)
    from oracle import decoder_oracle as do
    from oracle import encoder_oracle as eo
    esd = {k: v.to(dev) for k, v in eo.random_encoder_state(seed=0, layer_scale=1.0).items()}
    dsd = {k: v.to(dev) for k, v in do.random_transformer_decoder_state(0, V, end_bias=3.2).items()}
    imgs = synthetic_images(3, 5).to(dev)
    with torch.no_grad():
        for i in range(3):
            if i == 1:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            feats = eo.encoder_forward(esd, imgs[i:i + 1], 7)
            do.beam_search(dsd, feats, "transformer", 5, V - 2, V - 1, V, max_steps=50)
        torch.cuda.synchronize()
        sec = (time.perf_counter() - t0) / 2
    res["beam_search_transformer_k5"] = {"captions_per_sec": 1.0 / sec, "seconds_per_caption": sec,
                                         "how": "caption.py:160-255 semantics (one image at a time, beams as the "
                                                "batch, prefix re-run every step), fp32",
                                         "x_over_torch_eager": ours["beam"] * sec}
    return res


def _torch_eager_encoder(dev, B=64):
    import torchvision
    net = torchvision.models.convnext_base(weights=None).features.to(dev).eval()
    pool = torch.nn.AdaptiveAvgPool2d((7, 7))
    x = torch.randn(B, 3, 256, 256, device=dev)
    res = {"batch": B, "what": "torchvision convnext_base().features + AdaptiveAvgPool2d(7) + permute, eval, "
                               "torch " + torch.__version__ + " eager kernels (cuDNN/cuBLAS), random weights"}

    def run(fn):
        with torch.no_grad():
            ms = _timed(fn, 10, 3, dev, 1)
        return {"images_per_sec": B / (ms * 1e-3), "ms_per_step": ms}

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    res["fp32"] = run(lambda: pool(net(x)).permute(0, 2, 3, 1))
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    res["tf32"] = run(lambda: pool(net(x)).permute(0, 2, 3, 1))
    xc = x.contiguous(memory_format=torch.channels_last)
    netc = net.to(memory_format=torch.channels_last)

    def amp():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return pool(netc(xc)).permute(0, 2, 3, 1)
    res["bf16_autocast_channels_last"] = run(amp)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return res


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "encoder"],
                    help="headline line: the train step (BASELINE.json configs[3]) or Encoder.forward (configs[1])")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"], help="--workload encoder only")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary train / beam-search workloads")
    ap.add_argument("--eager-step", action="store_true",
                    help="headline train step through caption_train_step + DistributedDataParallel (eager launches) "
                         "instead of CapturedTrainStep (CUDA-graph replay, own gradient buckets)")
    ap.add_argument("--spans", default=None, help="write the per-launch timing table of the instrumented pass here")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # stdout carries exactly ONE JSON line: anything libraries print (e.g. "NCCL version ...") goes to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
