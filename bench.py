#!/usr/bin/env python
"""Headline benchmark: dMel encode audio-seconds per second (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of BASELINE configs[1]:
64 utterances x 10 s, 24 kHz, n_fft 1024, hop 256, 128 mel, 16 bins —
encode (waveform -> uint8 codes) and dequantise (codes -> mel) in ONE fused launch (the quantiser's
forward, ``DMelTokenizer.encode_decode``).  The K-step window is timed between two CUDA events and repeated
(``repetitions`` in the line, at least 30); ``value`` comes from the MEDIAN window, so one straggling window of
a 2 ms measurement cannot set the number.  A separate pass times single launches for the roofline entries.
With N GPUs every rank runs that batch on its own shard of utterances (weak scaling, no data-path collective).

After the headline region the default run measures, under ``secondary``, every other BASELINE config through
the same library: configs[2] (dataset calibrate + encode, sharded, the NCCL all-reduce INSIDE the timed region),
configs[3] (streaming, per-chunk latency), configs[4] (long-form n_fft 2048), and the stand-alone quantiser
kernels at a size where HBM binds.  ``torch_gpu_baseline`` is the reference's own file on CUDA tensors on this
GPU (the stock-op chain the fused kernel replaces: cuFFT + cuBLAS + elementwise kernels).

``--impl reference`` runs the reference's own, unmodified ``dmel_codec/utils/spectrogram.py`` (pip-installed
into ``baseline/_ref`` by ``__graft_entry__.build()``; ``librosa.filters.mel``, which cannot be installed
offline, is stood in for) on all host cores, followed by the quantiser spec in plain torch ops — the reference
has no quantiser.  If ``baseline/_ref`` is absent it falls back to the oracle port and says so (``kind``).
"""
from __future__ import annotations

import argparse
import collections
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

# ---- workload: BASELINE.json configs[1] -------------------------------------
SAMPLE_RATE = 24000
SECONDS = 10
BATCH = 64
N_BINS = 16
GEOM = dict(sample_rate=SAMPLE_RATE, n_fft=1024, win_length=1024, hop_length=256, n_mels=128, f_min=0.0, f_max=12000.0)
N_SAMPLES = SAMPLE_RATE * SECONDS
N_FRAMES = N_SAMPLES // GEOM["hop_length"]
AUDIO_SEC_PER_BATCH = BATCH * SECONDS
ENCODE_BYTES = 4 * BATCH * N_SAMPLES + BATCH * GEOM["n_mels"] * N_FRAMES
DEQUANT_BYTES = 5 * BATCH * GEOM["n_mels"] * N_FRAMES
# encode + dequant in one launch: the waveform in, codes and float32 bin centres out; the codes are not re-read
FORWARD_BYTES = 4 * BATCH * N_SAMPLES + 5 * BATCH * GEOM["n_mels"] * N_FRAMES
RING = 4  # distinct input batches cycled through so no step finds its input in the 126 MB L2
METRIC = "dmel_encode_audio_seconds_per_second"
UNIT = "audio-s/s"
WORKLOAD = ("configs[1]: 24 kHz speech, 128 mel, 16 bins, batch 64x10 s, n_fft 1024, hop 256; "
            "step = one fused launch: wav -> u8 codes + dequantised mel (bin centres)")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        if os.environ.get("DMEL_BENCH_NO_CLOCKS"):  # debugging aid: no sampler process at all
            return self
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            # nvidia-smi's start-up (NVML initialisation, device enumeration) holds the driver for 0.1 - 2 s, once 24 s,
            # and stalls whatever kernel sequence is running: wait for its first sample before the timed region begins
            deadline = time.perf_counter() + 30.0
            while not self.rows and self.proc.poll() is None and time.perf_counter() < deadline:
                time.sleep(0.01)
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, r[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def config_block():
    """Identical in both arms, so the driver's same-config check compares like with like."""
    return {"workload": WORKLOAD, "per_gpu_batch": f"{BATCH}x{SECONDS}s", "sample_rate": SAMPLE_RATE,
            "n_fft": GEOM["n_fft"], "hop": GEOM["hop_length"], "n_mels": GEOM["n_mels"], "n_bins": N_BINS}


def torch_quantizer_forward(mel, lo, scale, step, n_bins):
    """The quantiser spec in stock torch ops: what a user without this library would
    write after the reference's mel transform.  Used by the two baselines only, never by the product."""
    codes = torch.clamp(torch.floor((mel - lo[None, :, None]) * scale[None, :, None]), 0, n_bins - 1).to(torch.uint8)
    return codes, lo[None, :, None] + (codes.to(torch.float32) + 0.5) * step[None, :, None]


def torch_stats(mel, n_bins):
    lo, hi = mel.amin(dim=(0, 2)), mel.amax(dim=(0, 2))
    width = hi - lo
    scale = torch.where(width > 0, n_bins / width, torch.zeros_like(width))
    return lo, scale, width / n_bins


def reference_transform(mel_fn):
    """(transform, kind): the reference's own LogMelSpectrogram from baseline/_ref, or None when it is absent."""
    from baseline import ref_arm
    if not ref_arm.available():
        return None
    mod = ref_arm.load(mel_fn)
    return mod.LogMelSpectrogram(sample_rate=SAMPLE_RATE, n_fft=GEOM["n_fft"], win_length=GEOM["win_length"],
                                 hop_length=GEOM["hop_length"], n_mels=GEOM["n_mels"], f_min=GEOM["f_min"],
                                 f_max=GEOM["f_max"])


def cpu_reference_runner():
    """-> (step(wav) callable, setup(rows) -> wav, kind, description).  The reference arm's step: the reference's
    own transform (unmodified file) + the quantiser spec in torch ops; the oracle port only if baseline/_ref is
    missing."""
    from dmel_codec_b200 import synth
    from oracle import dmel_oracle as O  # test infrastructure; allowed in the reference arm / cpu_baseline leg only

    ref = reference_transform(O.slaney_filterbank)
    if ref is not None:
        state = {}

        def step(wav):
            with torch.no_grad():
                mel = ref(wav)
                if not state:
                    state["stats"] = torch_stats(mel[:4], N_BINS)
                return torch_quantizer_forward(mel, *state["stats"], N_BINS)

        kind = "reference"
        what = ("reference dmel_codec/utils/spectrogram.py (unmodified, baseline/_ref) + Appendix-B quantiser in torch ops; "
                "librosa.filters.mel stood in for")
    else:
        cfg = O.MelConfig(**GEOM)
        bank = torch.from_numpy(O.slaney_filterbank(SAMPLE_RATE, GEOM["n_fft"], GEOM["n_mels"], GEOM["f_min"], GEOM["f_max"]))
        state = {}

        def step(wav):
            mel = O.log_mel(wav, cfg, bank)
            if not state:
                state["stats"] = O.calibrate_minmax(mel[:4])
            lo, hi = state["stats"]
            codes = O.dmel_encode(mel, lo, hi, N_BINS)
            return codes, O.dmel_decode(codes, lo, hi, N_BINS)

        kind = "port"
        what = "oracle/dmel_oracle.py port of the reference's torch path (baseline/_ref absent)"

    def setup(rows):
        return synth.batch(range(rows), N_SAMPLES, SAMPLE_RATE, "speech")

    return step, setup, kind, what


def run_reference(args, rank: int):
    """The reference's CPU path on the host cores, same batch (rank 0 only; other ranks exit without work)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, setup, kind, what = cpu_reference_runner()
    wav = setup(BATCH)
    steps = min(args.steps, 100)  # each step is a full 64 x 10 s batch on the CPU (~0.1 s): bounded
    for _ in range(max(1, min(args.warmup, 3))):
        step(wav)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(wav)
    dt = time.perf_counter() - t0
    value = AUDIO_SEC_PER_BATCH * steps / dt
    sample = f"{BATCH} x {SECONDS} s per step, {steps} steps, {what}; torch {torch.__version__} CPU, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_block(),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def pin_to_gpu_numa_node(local_rank: int):
    """Run this rank on the CPUs next to its GPU (sysfs local_cpulist of the GPU's PCI device) so the pinned
    host buffers of the e2e leg are allocated on, and copied from, the memory of that NUMA node."""
    if os.environ.get("DMEL_BENCH_NO_AFFINITY"):
        return None
    try:
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id  # torch >= 2.x
    except Exception:
        try:
            import pynvml
            pynvml.nvmlInit()
            bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local_rank)).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
        except Exception:
            return None
    try:
        bus = str(bus).lower()
        if len(bus.split(":")[0]) == 8:  # nvml prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return spec
    except (OSError, ValueError):
        pass
    return None


REPETITIONS = 30  # K-step windows timed back to back; the headline is the median window


def median_window_ms(run_window, reps: int, stream) -> list:
    """Device time of `reps` windows, each between its own pair of CUDA events (events only BETWEEN windows, so a
    window is an uninterrupted kernel sequence)."""
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    queued = [time.perf_counter()]
    marks[0].record(stream)
    for r in range(reps):
        run_window(r)
        marks[r + 1].record(stream)
        queued.append(time.perf_counter())
    torch.cuda.synchronize()
    if os.environ.get("DMEL_BENCH_DEBUG"):
        print("host ms to queue each window: " + " ".join(f"{(b - a) * 1e3:.1f}" for a, b in zip(queued, queued[1:])), file=sys.stderr)
    return [marks[r].elapsed_time(marks[r + 1]) for r in range(reps)]


def time_kernel_ms(fn, reps: int, stream, keep: int = 0) -> float:
    """Median device time of one call of fn(i): every call sits between its own pair of events, the calls are queued
    back to back and the host synchronises once at the end, so the GPU never idles between them (a launch timed from
    an idle GPU carries ~15 us of ramp that no pipeline ever sees)."""
    marks = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    # keep > 0: the outputs of the last `keep` calls stay alive, so small outputs rotate through that many buffers instead
    # of landing on the same lines of L2 (allocated by untimed calls first)
    if keep:  # put the rotating buffers into the allocator's cache first: no cudaMalloc inside an event pair
        warm = [fn(i) for i in range(keep + 1)]
        del warm
    recent = collections.deque(maxlen=max(keep, 1))
    for i, (e0, e1) in enumerate(marks):
        e0.record(stream)
        recent.append(fn(i))
        e1.record(stream)
    torch.cuda.synchronize()
    return statistics.median(e0.elapsed_time(e1) for e0, e1 in marks)


def run_ours(args, rank: int, world: int, local_rank: int):
    import torch.distributed as dist
    import dmel_codec_b200 as d
    from dmel_codec_b200 import filters, synth
    from dmel_codec_b200 import plan as P

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: dmel_codec_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    affinity = pin_to_gpu_numa_node(local_rank) if world > 1 else None
    tok = d.DMelTokenizer(n_bins=N_BINS, **GEOM).to(dev)

    # this rank's shard: utterance ids [rank*RING*BATCH, ...): RING distinct batches
    base = rank * RING * BATCH
    ring = [synth.device_batch(range(base + r * BATCH, base + (r + 1) * BATCH), N_SAMPLES, SAMPLE_RATE, dev)
            for r in range(RING)]
    # dataset-wide calibration: local min/max then the one all-reduce of the path
    tok.calibrate(ring)
    q = tok.quantizer
    plan = tok._plan(dev)
    lo, scale, table, width = q.lo, q.scale(), q.table(), q.step()
    launch_cfg = plan.describe()
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream(dev)

    def step(i):  # the benchmark step: codes and dequantised mel from one launch
        return plan.encode_decode(ring[i % RING], None, lo, scale, width, N_BINS)

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # per-launch durations (roofline): a separate pass with an event on either side of each kernel
    probe = max(50, min(args.steps, 200))  # 5 ms of launches: enough for a median that the ramp of the first few does not move
    fwd_ms = time_kernel_ms(lambda i: plan.encode_decode(ring[i % RING], None, lo, scale, width, N_BINS), probe, stream, keep=8)
    enc_ms = time_kernel_ms(lambda i: plan.encode(ring[i % RING], None, lo, scale, N_BINS), probe, stream, keep=8)
    codes0 = plan.encode(ring[0], None, lo, scale, N_BINS)
    deq_ms = time_kernel_ms(lambda i: P.dequantize(codes0, table), probe, stream, keep=8)
    if world > 1:
        dist.barrier()

    # ---- the timed region: REPETITIONS windows of exactly K steps, barrier + synchronize on both sides ----------
    reps = max(REPETITIONS, 1)
    with ClockSampler(local_rank) as clocks:
        torch.cuda.synchronize()
        wall0 = time.perf_counter()
        # the outputs of the last 8 steps stay alive, so the allocator hands every step a buffer that was last written
        # eight steps (8 x 38 MB > L2) ago; keeping ALL of a window's outputs alive, as a list comprehension does, makes
        # the first window allocate K x 38 MB of fresh HBM (7 - 24 s of cudaMalloc at K = 2000)
        recent = collections.deque(maxlen=8)

        def window(r):
            for i in range(args.steps):
                recent.append(step(args.warmup + r * args.steps + i))

        windows = median_window_ms(window, reps, stream)
        recent.clear()
        wall = time.perf_counter() - wall0
    if os.environ.get("DMEL_BENCH_DEBUG"):
        print(f"[rank {rank}] windows (ms): " + " ".join(f"{w:.3f}" for w in windows), file=sys.stderr)
    if world > 1:
        dist.barrier()
    t = torch.tensor([statistics.median(windows), min(windows), max(windows), wall * 1e3 / reps, fwd_ms, enc_ms, deq_ms],
                     dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # every figure is the slowest rank's
    total_ms, win_min, win_max, wall_ms, fwd_avg_ms, enc_avg_ms, deq_avg_ms = t.tolist()

    # ---- end to end through the public API with HOST buffers -------------------
    e2e = e2e_pcm = None
    h2d_bytes, d2h_bytes = 4 * BATCH * N_SAMPLES, BATCH * GEOM["n_mels"] * N_FRAMES
    if not args.skip_e2e:
        e2e_steps = max(3, min(args.steps, 100))  # ~1.3 ms each: long enough to average out host jitter
        host = [ring[r].cpu().pin_memory() for r in range(2)]
        out = torch.empty((BATCH, GEOM["n_mels"], N_FRAMES), dtype=torch.uint8).pin_memory()

        def timed_calls(bufs):
            for r in range(2):
                tok.encode_host(bufs[r], out=out)
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for i in range(e2e_steps):
                tok.encode_host(bufs[i % 2], out=out)  # returns after the codes are in host memory
            e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(e, op=dist.ReduceOp.MAX)
            return e.item() / e2e_steps

        # the raw pinned host->device rate of this box in this run: the floor of any host-buffer call
        # (best of 12 single copies of the batch, each between its own events)
        scratch = torch.empty_like(ring[0])
        best = float("inf")
        for i in range(14):
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(stream)
            scratch.copy_(host[i % 2], non_blocking=True)
            c1.record(stream)
            c1.synchronize()
            if i >= 2:
                best = min(best, c0.elapsed_time(c1))
        h2d_peak = h2d_bytes / (best / 1e3) / 1e9
        del scratch
        dt = timed_calls(host)
        e2e = {"value": world * AUDIO_SEC_PER_BATCH / dt, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
               "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps, "ms_per_call": dt * 1e3,
               "h2d_gbs": h2d_bytes / dt / 1e9, "h2d_peak_gbs": h2d_peak,
               "frac_of_h2d_floor": (h2d_bytes / (h2d_peak * 1e9)) / dt,
               "call": "DMelTokenizer.encode_host -> dmel_encode_host_u8 (pinned host wav in, host codes out); "
                       "h2d_peak_gbs = raw pinned cudaMemcpyAsync of the same bytes, same run, this rank"}
        # the same call with the waveform as int16 PCM (the format audio is stored in): half the H2D bytes
        host_pcm = [(h.clamp(-1, 1) * 32767.0).round().to(torch.int16).pin_memory() for h in host]
        dt = timed_calls(host_pcm)
        e2e_pcm = {"value": world * AUDIO_SEC_PER_BATCH / dt, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes // 2,
                   "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps, "ms_per_call": dt * 1e3,
                   "call": "same call with int16 PCM host waveforms (dmel_encode_host_pcm16_u8); secondary line, "
                           "the reference interface takes float32"}
        del host, host_pcm, out

    # ---- the stock GPU implementation: the reference's own file on CUDA tensors, same GPU, same batch ----------
    torch_gpu = None
    if not args.skip_gpu_baseline:
        ref = reference_transform(filters.mel_filterbank)
        if ref is not None:
            with torch.no_grad():
                stats = torch_stats(ref(ring[0][:4]), N_BINS)

                def ref_step(i):
                    return torch_quantizer_forward(ref(ring[i % RING]), *stats, N_BINS)

                for i in range(3):
                    ref_step(i)
                k = max(3, min(args.steps, 50))
                def ref_window(r):
                    for i in range(k):
                        ref_step(r * k + i)

                ms = statistics.median(median_window_ms(ref_window, 5, stream)) / k
            g = torch.tensor([ms], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(g, op=dist.ReduceOp.MAX)
            torch_gpu = {"value": world * AUDIO_SEC_PER_BATCH / (g.item() / 1e3), "unit": UNIT, "ms_per_step": g.item(),
                         "what": "reference dmel_codec/utils/spectrogram.py (unmodified, baseline/_ref) on CUDA tensors: reflect pad, "
                                 "torch.stft (cuFFT), magnitude, matmul (cuBLAS), log-clamp, then the quantiser spec in torch ops; "
                                 "same batch ring, device-resident, median of 5 windows"}
        else:
            torch_gpu = {"value": None, "unavailable": "baseline/_ref absent: run __graft_entry__.build() where /root/reference exists"}

    # ---- the other BASELINE configs, through the same library --------------------------------------------------
    secondary = None
    if not args.skip_secondary:
        del ring
        torch.cuda.empty_cache()
        secondary = {}
        secondary["configs[2]"] = pool_workload("calibrate", rank, world, dev)
        secondary["configs[2] variable length"] = pool_workload("calibrate_varlen", rank, world, dev)
        secondary["configs[4]"] = pool_workload("longform", rank, world, dev)
        if rank == 0:
            secondary["configs[3]"] = stream_workload(dev)
            secondary["kernels"] = standalone_kernels(dev)
        if world > 1:
            dist.barrier()

    if rank != 0:
        return
    peak, peak_src = measured_peaks()
    enc_gbs = ENCODE_BYTES / (enc_avg_ms / 1e3) / 1e9
    deq_gbs = DEQUANT_BYTES / (deq_avg_ms / 1e3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "encode_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source", "ncu --set full capture, profiles/encode_traffic.json (not measured in this run)")

    # CPU baseline: the reference arm's step on this box's host cores, bounded sample
    cores = os.cpu_count() or 1
    cpu = {"value": None, "unit": UNIT, "cores": cores, "kind": "reference", "sample": "skipped"}
    if not args.skip_cpu and world == 1:  # the CPU baseline is a 1-GPU-run line (rank 0, N = 1 only)
        torch.set_num_threads(cores)
        cstep, csetup, kind, what = cpu_reference_runner()
        cwav = csetup(BATCH)
        cstep(cwav)
        passes, c0 = 0, time.perf_counter()
        while passes < 3 or (time.perf_counter() - c0 < 10.0 and passes < 200):
            cstep(cwav)
            passes += 1
        cpu_dt = time.perf_counter() - c0
        cpu = {"value": AUDIO_SEC_PER_BATCH * passes / cpu_dt, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{passes} passes of the same {BATCH}x{SECONDS}s batch in {cpu_dt:.1f} s: {what}; {cores} threads"}

    value = world * AUDIO_SEC_PER_BATCH * args.steps / (total_ms / 1e3)
    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config_block(),
        "timing": {"repetitions": reps, "launches_per_window": args.steps, "window_ms_median": total_ms, "window_ms_min": win_min, "window_ms_max": win_max,
                   "wall_ms_per_window": wall_ms, "sharding": "utterances, no data-path collective",
                   "l2": f"inputs cycle through a ring of {RING} distinct batches ({RING * ENCODE_BYTES / 1e6:.0f} MB > 126 MB L2)",
                   "method": "each window = K steps between two CUDA events, no events inside; value from the median window, "
                             "max over ranks; per-launch durations from a separate pass",
                   "rank0_cpu_affinity": affinity},
        "roofline": {"bound": "hbm", "kernel": f"dmel_fused_kernel<{GEOM['n_fft']},{launch_cfg['tile_frames']},codes+dequant> "
                     f"({launch_cfg['ctas_per_sm']} CTA/SM, {launch_cfg['smem_bytes']} B smem)",
                     "achieved": FORWARD_BYTES / (fwd_avg_ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": FORWARD_BYTES / (fwd_avg_ms / 1e3) / 1e9 / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": FORWARD_BYTES, "avg_launch_ms": fwd_avg_ms,
                     "note": "the step's kernel, one launch between two events (probe pass, median). Algorithmic bytes = 4*B*L waveform in "
                             "+ B*M*T codes out + 4*B*M*T dequantised mel out (SURVEY 8d encode + dequant, minus the code re-read)",
                     "encode_only": {"achieved": enc_gbs, "frac": enc_gbs / peak, "algorithmic_bytes_per_launch": ENCODE_BYTES,
                                     "avg_launch_ms": enc_avg_ms},
                     "dequant": {"achieved": deq_gbs, "frac": deq_gbs / peak, "algorithmic_bytes_per_launch": DEQUANT_BYTES,
                                 "avg_launch_ms": deq_avg_ms,
                                 "note": "38 MB launch: latency-sized; secondary.kernels has it at a size where HBM binds"}},
        "cpu_baseline": cpu,
        "torch_gpu_baseline": torch_gpu,
        "e2e": e2e if e2e is not None else {"value": None, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
        "e2e_pcm16": e2e_pcm,
        "secondary": secondary,
        "gpu_launches": args.steps * reps,  # every step is one launch of dmel_fused_kernel; `reps` windows of K steps were timed
        "clocks": clocks.summary(),
    }))


# ---------------------------------------------------------------------------
# secondary workloads (BASELINE configs[2], [3], [4]) and stand-alone kernels
# ---------------------------------------------------------------------------
def _init_rank():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    return rank, world, local_rank, torch.device("cuda", local_rank)


def stream_workload(dev):
    """configs[3]: one stream, 80 ms chunks, 16 kHz / 80 mel: per-chunk latency (host call + stream sync)."""
    import dmel_codec_b200 as d
    from dmel_codec_b200 import synth
    geom = dict(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=256, n_mels=80)
    tok = d.DMelTokenizer(n_bins=16, **geom).to(dev)
    wav = synth.batch([0], 16000 * 30, 16000, "speech").to(dev)
    tok.quantizer.reset_stats()
    tok.update_stats(wav)  # local statistics only: this workload runs on rank 0 alone, so NO collective may be called here
    enc = d.DMelStreamEncoder(tok, n_streams=1, capacity_samples=1 << 16)
    chunk, lat = 1280, []
    n_chunks = wav.shape[2] // chunk
    outs = {}  # a streaming server reuses its output buffers: one per distinct frame count (3, then 5 per chunk)
    for rep in range(3):  # first pass warms up
        lat = []
        for i in range(n_chunks):
            x = wav[:, 0, i * chunk:(i + 1) * chunk]
            k = enc.frames_after(chunk)
            if k not in outs:
                outs[k] = torch.empty((1, geom["n_mels"], k), dtype=torch.uint8, device=dev)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            enc.push(x, out=outs[k])
            torch.cuda.synchronize()
            lat.append((time.perf_counter() - t0) * 1e6)
        enc.flush()
    lat.sort()
    p50, p99 = lat[len(lat) // 2], lat[min(len(lat) - 1, int(len(lat) * 0.99))]
    # zero-copy form: the producer has written the chunk straight into the history buffer (input_view); the timed
    # part is commit() + stream sync, i.e. one kernel launch
    zc = []
    for rep in range(3):
        zc = []
        for i in range(n_chunks):
            k = enc.frames_after(chunk)
            enc.input_view(chunk).copy_(wav[:, 0, i * chunk:(i + 1) * chunk])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            enc.commit(chunk, out=outs[k])
            torch.cuda.synchronize()
            zc.append((time.perf_counter() - t0) * 1e6)
        enc.flush()
    zc.sort()
    # the floor of ANY launch-and-wait on this box: an empty torch kernel + stream sync
    tiny, floor = torch.zeros(8, device=dev), []
    for _ in range(300):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tiny.add_(1)
        torch.cuda.synchronize()
        floor.append((time.perf_counter() - t0) * 1e6)
    floor.sort()
    return {"metric": "dmel_stream_chunk_latency_us", "value": p50, "unit": "us (p50)", "p99_us": p99,
            "zero_copy_p50_us": zc[len(zc) // 2], "zero_copy_p99_us": zc[min(len(zc) - 1, int(len(zc) * 0.99))],
            "launch_and_sync_floor_p50_us": floor[len(floor) // 2],
            "higher_is_better": False, "n_gpus": 1, "chunks": len(lat),
            "audio_seconds_per_second_one_stream": 0.08 / (sum(lat) / len(lat) * 1e-6),
            "config": {"workload": "configs[3]: batch 1, 80 ms chunks (1280 samples), 16 kHz, 80 mel, 16 bins; "
                                   "chunk already on the device, latency = push() + stream sync; 5 frames per chunk",
                       "note": "launch-latency bound: 5.5 KB per call, byte roofline not meaningful. value = push() (chunk copy + "
                               "launch); zero_copy = input_view()/commit() (launch only); floor = an empty kernel + sync on this box"}}


def pool_workload(name, rank, world, dev):
    """configs[2] (calibrate + encode of 10k utterances, sharded) and configs[4] (long-form 2048).  The timed
    region is the whole job of this rank's shard including the statistics all-reduce (NCCL when world > 1);
    max over ranks."""
    import torch.distributed as dist
    import dmel_codec_b200 as d
    from dmel_codec_b200 import distributed as D, synth
    varlen = name == "calibrate_varlen"
    if name in ("calibrate", "calibrate_varlen"):
        geom = dict(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=256, n_mels=80)
        n_utts, seconds, n_bins, bsz, pool_n = 10000, (15 if varlen else 10), 16, 250, 500
        label = ("configs[2], variable-length variant (SURVEY 8d): 10,000 utterances of 2-15 s (uniform, seeded), right-padded to 15 s "
                 "batches with audio_lengths; frames past each length are neither calibrated nor encoded" if varlen else
                 "configs[2]: min/max calibration + encode of 10,000 x 10 s utterances (16 kHz, 80 mel, 16 bins), block-sharded")
    else:
        geom = dict(sample_rate=44100, n_fft=2048, win_length=2048, hop_length=512, n_mels=160)
        n_utts, seconds, n_bins, bsz, pool_n = 32, 60, 32, 4, 4
        label = "configs[4]: long-form 32 x 60 s, 44.1 kHz, n_fft 2048, hop 512, 160 mel, 32 bins, utterances sharded"
    n = geom["sample_rate"] * seconds
    tok = d.DMelTokenizer(n_bins=n_bins, **geom).to(dev)
    # a pool of distinct synthetic utterances stands in for the shard (content does not change the work)
    pool = synth.device_batch(range(rank * pool_n, (rank + 1) * pool_n), n, geom["sample_rate"], dev)

    lengths_all = None
    if varlen:  # one seeded length per utterance id, the same on every rank
        g = torch.Generator().manual_seed(20261018)
        lengths_all = torch.randint(2 * geom["sample_rate"], n + 1, (n_utts,), generator=g, dtype=torch.int32).to(dev)

    def load(ids):
        k = len(ids)
        start = (ids[0] * 7) % max(1, pool_n - k + 1) if k <= pool_n else 0
        if varlen:
            return pool[start:start + k], lengths_all[ids[0]:ids[0] + k]
        return pool[start:start + k]

    two_pass = bool(os.environ.get("DMEL_BENCH_TWO_PASS"))

    def job():
        if two_pass:
            D.calibrate_sharded(tok, n_utts, load, bsz)            # pass 1 + the all-reduce
            for _ in D.encode_sharded(tok, n_utts, load, bsz):     # pass 2: the transform again
                pass
        else:  # pass 1 keeps the shard's log-mel in HBM, pass 2 is the stand-alone quantiser over it
            for _ in D.calibrate_encode_sharded(tok, n_utts, load, bsz):
                pass

    job()
    times = []
    for _ in range(9):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        job()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        times.append(ms.item())
    ms_total = statistics.median(times)
    audio_s = float(lengths_all.sum().item()) / geom["sample_rate"] if varlen else n_utts * seconds
    t_frames = n // geom["hop_length"]
    bytes_alg = 2 * 4 * n_utts * n + n_utts * geom["n_mels"] * t_frames  # two waveform reads + codes
    bytes_job = 4 * n_utts * n + (4 + 4 + 1) * n_utts * geom["n_mels"] * t_frames  # what this job moves: wav, mel out, mel in, codes
    peak, _ = measured_peaks()
    out = {"metric": f"dmel_{name}_encode_audio_seconds_per_second", "value": audio_s / (ms_total / 1e3), "unit": UNIT,
           "audio_seconds": audio_s,
           "n_gpus": world, "ms_total": ms_total, "ms_runs": times, "higher_is_better": True, "scaling": "strong",
           "hbm_frac_all_gpus": bytes_alg / (ms_total / 1e3) / 1e9 / (peak * world),
           "hbm_frac_bytes_moved": bytes_job / (ms_total / 1e3) / 1e9 / (peak * world),
           "collective": "all_reduce(MIN) of [lo, -hi] (2 * n_mels float32) inside the timed region" + ("" if world > 1 else " (no-op at 1 rank)"),
           "stats": {"lo_min": float(tok.quantizer.lo.min()), "hi_max": float(tok.quantizer.hi.max())},
           "config": {"workload": label, "launch": tok._plan(dev).describe(),
                      "job": "two transform passes" if two_pass else
                             "one transform pass (log-mel of the shard kept in HBM) + stand-alone quantiser pass",
                      "data": f"pool of {pool_n} distinct synthetic utterances per rank, cycled; median of 9 runs"}}
    del pool, tok
    torch.cuda.empty_cache()
    return out


def standalone_kernels(dev):
    """Roofline entries of the stand-alone quantiser stages at a size where HBM binds (1 GiB of log-mel: 8x the L2)."""
    from dmel_codec_b200 import plan as P
    b, m, t = 256, 128, 8192
    mel = torch.empty((b, m, t), dtype=torch.float32, device=dev).uniform_(-11.5, 2.0)
    lo = torch.full((m,), -11.6, dtype=torch.float32, device=dev)
    hi = torch.full((m,), 2.1, dtype=torch.float32, device=dev)
    scale, step = 16.0 / (hi - lo), (hi - lo) / 16.0
    table = (lo[:, None] + (torch.arange(16, device=dev, dtype=torch.float32)[None, :] + 0.5) * step[:, None]).contiguous()
    run_min = torch.full((m,), float("inf"), device=dev)
    run_max = torch.full((m,), float("-inf"), device=dev)
    stream = torch.cuda.current_stream(dev)
    codes = P.quantize(mel, lo, scale, 16)
    peak, _ = measured_peaks()
    n = b * m * t
    out = {}
    from dmel_codec_b200 import FSQIndexer
    fsq = FSQIndexer(levels=(7, 5, 5), groups=10)              # reference config/lm/lm_config.yaml:95-107
    zb, zt = 64, 65536
    zp = torch.randn((zb, zt, 10, 3), dtype=torch.float32, device=dev)
    n_fsq = zb * zt * 10
    gain_wav = torch.empty((64, 1 << 21), dtype=torch.float32, device=dev).uniform_(-0.5, 0.5)  # 512 MiB of waveform
    plan = __import__("dmel_codec_b200").LogMelSpectrogram(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=256,
                                                          n_mels=80).spectrogram.plan_for(dev)
    for name, fn, nbytes in (
            ("quantize_kernel", lambda i: P.quantize(mel, lo, scale, 16), 5 * n),
            ("dequantize_kernel", lambda i: P.dequantize(codes, table), 5 * n),
            ("tensor_minmax_kernel", lambda i: P.tensor_minmax(mel, None, run_min, run_max), 4 * n),
            ("row_absmax_kernel (peak normalisation gain)", lambda i: plan.peak_gain(gain_wav), 4 * gain_wav.numel()),
            ("fsq_encode_kernel (indices + lm ids)", lambda i: fsq.encode(zp, return_codes=False, lm_codebook_size=180), (12 + 16) * n_fsq)):
        for i in range(2):
            fn(i)
        ms = time_kernel_ms(fn, 7, stream)
        out[name] = {"avg_launch_ms": ms, "algorithmic_bytes_per_launch": nbytes, "achieved": nbytes / (ms / 1e3) / 1e9,
                     "unit": "GB/s", "frac": nbytes / (ms / 1e3) / 1e9 / peak}
    del mel, codes, zp, gain_wav
    torch.cuda.empty_cache()
    out["antialias_snake_kernel"] = activation_entry(dev, peak)
    return out


def activation_entry(dev, peak):
    """BigVGAN's anti-aliased Snake activation: this library's one-pass kernel against the reference's
    own torch path (UpSample1d -> SnakeBeta -> DownSample1d) on the same GPU; 4 bytes in + 4 bytes out per sample."""
    import dmel_codec_b200 as d
    b, c, t = 8, 256, 32768
    x = torch.randn((b, c, t), dtype=torch.float32, device=dev)
    mod = d.AntiAliasSnake(c).to(dev)
    mod.alpha.normal_(0.0, 0.5)
    mod.beta.normal_(0.0, 0.5)
    y = torch.empty_like(x)
    stream = torch.cuda.current_stream(dev)
    for _ in range(2):
        mod(x, out=y)
    ms = time_kernel_ms(lambda i: mod(x, out=y), 7, stream)
    nbytes = 8 * x.numel()
    entry = {"avg_launch_ms": ms, "algorithmic_bytes_per_launch": nbytes, "achieved": nbytes / (ms / 1e3) / 1e9, "unit": "GB/s",
             "frac": nbytes / (ms / 1e3) / 1e9 / peak, "shape": [b, c, t]}
    try:
        from baseline import ref_arm
        Activation1d, activations = ref_arm.load_activation()
        act = activations.SnakeBeta(c, alpha_logscale=True).to(dev)
        with torch.no_grad():
            act.alpha.copy_(mod.alpha)
            act.beta.copy_(mod.beta)
            ref = Activation1d(activation=act).to(dev)
            for _ in range(2):
                want = ref(x)
            ref_ms = time_kernel_ms(lambda i: ref(x), 5, stream)
        err = ((y - want).abs() / want.abs().clamp(min=1.0)).max().item()
        entry["reference_torch_path"] = {"avg_ms": ref_ms, "speedup": ref_ms / ms, "max_rel_err_vs_it": err,
                                         "what": "reference Activation1d(SnakeBeta) from baseline/_ref, unmodified, CUDA tensors, same GPU. "
                                                 "The reference's fused kernel is built for sm_70 / sm_80 without PTX and cannot run here"}
    except Exception as e:  # baseline/_ref absent
        entry["reference_torch_path"] = {"unavailable": str(e)[:200]}
    del x, y
    torch.cuda.empty_cache()
    return entry


def run_stream(args):
    rank, world, local_rank, dev = _init_rank()
    print(json.dumps(stream_workload(dev)))


def run_pool_workload(args, name):
    import torch.distributed as dist
    rank, world, local_rank, dev = _init_rank()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        import datetime
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev, timeout=datetime.timedelta(seconds=240))
    out = pool_workload(name, rank, world, dev)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=["encode", "stream", "calibrate", "longform"], default="encode",
                    help="encode = the headline configs[1] line; the others are secondary lines for DESIGN.md")
    ap.add_argument("--skip-cpu", action="store_true", help="profiling runs: leave out the CPU-baseline leg")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs: leave out the host-buffer leg")
    ap.add_argument("--skip-secondary", action="store_true", help="profiling runs: leave out configs[2..4] and the stand-alone kernels")
    ap.add_argument("--skip-gpu-baseline", action="store_true", help="profiling runs: leave out the stock-torch GPU comparator")
    ap.add_argument("--quick", action="store_true", help="headline region only (= all four --skip flags)")
    args = ap.parse_args()
    if args.quick:
        args.skip_cpu = args.skip_e2e = args.skip_secondary = args.skip_gpu_baseline = True
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.workload == "stream":
        return run_stream(args)
    if args.workload in ("calibrate", "longform"):
        return run_pool_workload(args, args.workload)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        torch.cuda.set_device(local_rank)
        import datetime
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(seconds=240))  # a mismatched collective fails in minutes, not in ten
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
