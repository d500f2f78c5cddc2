#!/usr/bin/env python
"""Headline benchmark: dMel encode audio-seconds per second (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of BASELINE configs[1]:
64 utterances x 10 s, 24 kHz, n_fft 1024, hop 256, 128 mel, 16 bins —
encode (waveform -> uint8 codes) and dequantise (codes -> mel) in ONE fused launch (the quantiser's
forward, ``DMelTokenizer.encode_decode``); a separate pass times the stand-alone encode and
dequantise kernels for the roofline entries.
With N GPUs every rank runs that batch on its own shard of utterances (weak
scaling, no data-path collective; the calibration all-reduce happens once,
before the timed region).  Prints ONE JSON line on rank 0.

``--impl reference`` times the CPU oracle port of the reference's torch path
(oracle/dmel_oracle.py; the reference is Python, so there is no oracle/_ref)
on the host cores for the same batch.
"""
from __future__ import annotations

import argparse
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
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
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


def oracle_step(wav, cfg, bank, lo, hi):
    from oracle import dmel_oracle as O
    codes = O.dmel_encode(O.log_mel(wav, cfg, bank), lo, hi, N_BINS)
    return O.dmel_decode(codes, lo, hi, N_BINS)


def cpu_oracle_setup(rows: int):
    from dmel_codec_b200 import synth
    from oracle import dmel_oracle as O
    cfg = O.MelConfig(**GEOM)
    bank = torch.from_numpy(O.slaney_filterbank(SAMPLE_RATE, GEOM["n_fft"], GEOM["n_mels"], GEOM["f_min"], GEOM["f_max"]))
    wav = synth.batch(range(rows), N_SAMPLES, SAMPLE_RATE, "speech")
    lo, hi = O.calibrate_minmax(O.log_mel(wav[: min(rows, 4)], cfg, bank))
    return wav, cfg, bank, lo, hi


def run_reference(args, rank: int):
    """The reference's CPU path (oracle port) on the host cores, same batch."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wav, cfg, bank, lo, hi = cpu_oracle_setup(BATCH)
    steps = min(args.steps, 100)  # each step is a full 64 x 10 s batch on the CPU (~0.1 s): bounded
    args.steps = steps
    for _ in range(min(args.warmup, 3)):
        oracle_step(wav, cfg, bank, lo, hi)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_step(wav, cfg, bank, lo, hi)
    dt = time.perf_counter() - t0
    value = AUDIO_SEC_PER_BATCH * args.steps / dt
    sample = f"{BATCH} x {SECONDS} s per step, {args.steps} steps, torch {torch.__version__} CPU ops, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU oracle port of reference utils/spectrogram.py + Appendix-B quantiser"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
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


def run_ours(args, rank: int, world: int, local_rank: int):
    import torch.distributed as dist
    import dmel_codec_b200 as d
    from dmel_codec_b200 import synth

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

    from dmel_codec_b200 import plan as P
    stream = torch.cuda.current_stream(dev)

    def step(i, ev=None):
        wav = ring[i % RING]
        if ev is None:  # the benchmark step: codes and dequantised mel from one launch
            return plan.encode_decode(wav, None, lo, scale, width, N_BINS)
        ev[0].record(stream)  # probe pass: the step's kernel and the two stand-alone kernels, each between events
        plan.encode_decode(wav, None, lo, scale, width, N_BINS)
        ev[1].record(stream)
        codes = plan.encode(wav, None, lo, scale, N_BINS)
        ev[2].record(stream)
        mel = P.dequantize(codes, table)
        ev[3].record(stream)
        return codes, mel

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # per-launch durations (roofline): a separate pass with an event on either side of each kernel, so the timed
    # region below is an uninterrupted kernel sequence, as in a real pipeline
    probe = max(3, min(args.steps, 200))
    events = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(probe)]
    for i in range(probe):
        step(i, events[i])
    torch.cuda.synchronize()
    fwd_ms = [e[0].elapsed_time(e[1]) for e in events]
    enc_ms = [e[1].elapsed_time(e[2]) for e in events]
    deq_ms = [e[2].elapsed_time(e[3]) for e in events]
    if world > 1:
        dist.barrier()
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        torch.cuda.synchronize()
        wall0 = time.perf_counter()
        t_begin.record(stream)
        for i in range(args.steps):
            step(args.warmup + i)
        t_end.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - wall0
    if world > 1:
        dist.barrier()
    total_ms = t_begin.elapsed_time(t_end)  # device time of the K back-to-back steps
    t = torch.tensor([total_ms, sum(enc_ms) / probe * args.steps, sum(deq_ms) / probe * args.steps, wall * 1e3,
                      sum(fwd_ms) / probe], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, enc_total, deq_total, wall_ms, fwd_avg_ms = t.tolist()

    # ---- end to end through the public API with HOST buffers -------------------
    e2e_steps = max(3, min(args.steps, 100))  # ~1.4 ms each: long enough to average out host jitter
    e2e_value = e2e_pcm_value = None
    if not args.skip_e2e:
        host = [ring[r].cpu().pin_memory() for r in range(2)]
        out = torch.empty((BATCH, GEOM["n_mels"], N_FRAMES), dtype=torch.uint8).pin_memory()
        for r in range(2):
            tok.encode_host(host[r], out=out)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            tok.encode_host(host[i % 2], out=out)  # returns after codes are in host memory
        e2e_dt = time.perf_counter() - t0
        e = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e, op=dist.ReduceOp.MAX)
        e2e_value = world * AUDIO_SEC_PER_BATCH * e2e_steps / e.item()
        # the same call with the waveform as int16 PCM (the format audio is stored in): half the H2D bytes
        host_pcm = [(h.clamp(-1, 1) * 32767.0).round().to(torch.int16).pin_memory() for h in host]
        for r in range(2):
            tok.encode_host(host_pcm[r], out=out)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            tok.encode_host(host_pcm[i % 2], out=out)
        e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e, op=dist.ReduceOp.MAX)
        e2e_pcm_value = world * AUDIO_SEC_PER_BATCH * e2e_steps / e.item()

    if rank != 0:
        return
    peak, peak_src = measured_peaks()
    enc_avg_s = enc_total / args.steps / 1e3
    achieved = ENCODE_BYTES / enc_avg_s / 1e9
    deq_gbs = DEQUANT_BYTES / (deq_total / args.steps / 1e3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "encode_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")

    # CPU baseline: the oracle port on this box's host cores, bounded sample
    cores = os.cpu_count() or 1
    cpu_value, passes, cpu_dt = None, 0, 0.0
    if not args.skip_cpu and world == 1:  # the CPU baseline is a 1-GPU-run line (rank 0, N = 1 only)
        torch.set_num_threads(cores)
        cwav, cfg, bank, clo, chi = cpu_oracle_setup(BATCH)
        oracle_step(cwav, cfg, bank, clo, chi)
        c0 = time.perf_counter()
        while passes < 3 or (time.perf_counter() - c0 < 10.0 and passes < 200):
            oracle_step(cwav, cfg, bank, clo, chi)
            passes += 1
        cpu_dt = time.perf_counter() - c0
        cpu_value = AUDIO_SEC_PER_BATCH * passes / cpu_dt

    value = world * AUDIO_SEC_PER_BATCH * args.steps / (total_ms / 1e3)
    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "per_gpu_batch": f"{BATCH}x{SECONDS}s", "sharding": "utterances, no data-path collective",
                   "l2": f"inputs cycle through a ring of {RING} distinct batches ({RING * ENCODE_BYTES / 1e6:.0f} MB > 126 MB L2)",
                   "timing": "K steps between two CUDA events, no events inside; per-launch durations from a separate pass",
                   "wall_ms_per_step": wall_ms / args.steps, "rank0_cpu_affinity": affinity},
        "roofline": {"bound": "hbm", "kernel": f"dmel_fused_kernel<{GEOM['n_fft']},{launch_cfg['tile_frames']},codes+dequant> "
                     f"({launch_cfg['ctas_per_sm']} CTA/SM, {launch_cfg['smem_bytes']} B smem)",
                     "achieved": FORWARD_BYTES / (fwd_avg_ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": FORWARD_BYTES / (fwd_avg_ms / 1e3) / 1e9 / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": FORWARD_BYTES, "avg_launch_ms": fwd_avg_ms,
                     "note": "the step's kernel, one launch between two events (probe pass). Algorithmic bytes = 4*B*L waveform in "
                             "+ B*M*T codes out + 4*B*M*T dequantised mel out (SURVEY 8d encode + dequant, minus the code re-read)",
                     "encode_only": {"achieved": achieved, "frac": achieved / peak, "algorithmic_bytes_per_launch": ENCODE_BYTES,
                                     "avg_launch_ms": enc_avg_s * 1e3},
                     "dequant": {"achieved": deq_gbs, "frac": deq_gbs / peak, "algorithmic_bytes_per_launch": DEQUANT_BYTES,
                                 "avg_launch_ms": deq_total / args.steps}},
        "cpu_baseline": {"value": cpu_value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{passes} passes of the same {BATCH}x{SECONDS}s batch, oracle/dmel_oracle.py, torch CPU, {cores} threads, {cpu_dt:.1f} s"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 4 * BATCH * N_SAMPLES,
                "d2h_bytes_per_step": BATCH * GEOM["n_mels"] * N_FRAMES, "steps": e2e_steps,
                "call": "DMelTokenizer.encode_host -> dmel_encode_host_u8 (pinned host wav in, host codes out)"},
        "e2e_pcm16": {"value": e2e_pcm_value, "unit": UNIT, "h2d_bytes_per_step": 2 * BATCH * N_SAMPLES,
                      "d2h_bytes_per_step": BATCH * GEOM["n_mels"] * N_FRAMES, "steps": e2e_steps,
                      "call": "same call with int16 PCM host waveforms (dmel_encode_host_pcm16_u8); secondary line, "
                              "the reference interface takes float32"},
        "gpu_launches": args.steps,
        "clocks": clocks.summary(),
    }))


# ---------------------------------------------------------------------------
# secondary workloads (BASELINE configs[2], [3], [4]); not the headline line
# ---------------------------------------------------------------------------
def _init_rank():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    return rank, world, local_rank, torch.device("cuda", local_rank)


def run_stream(args):
    """configs[3]: one stream, 80 ms chunks, 16 kHz / 80 mel: per-chunk latency."""
    import dmel_codec_b200 as d
    from dmel_codec_b200 import synth
    rank, world, local_rank, dev = _init_rank()
    geom = dict(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=256, n_mels=80)
    tok = d.DMelTokenizer(n_bins=16, **geom).to(dev)
    wav = synth.batch([0], 16000 * 30, 16000, "speech").to(dev)
    tok.calibrate([wav])
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
            codes = enc.push(x, out=outs[k])
            torch.cuda.synchronize()
            lat.append((time.perf_counter() - t0) * 1e6)
        enc.flush()
    lat.sort()
    p50, p99 = lat[len(lat) // 2], lat[min(len(lat) - 1, int(len(lat) * 0.99))]
    print(json.dumps({"metric": "dmel_stream_chunk_latency_us", "value": p50, "unit": "us (p50)", "p99_us": p99,
                      "higher_is_better": False, "n_gpus": 1, "chunks": len(lat),
                      "audio_seconds_per_second_one_stream": 0.08 / (sum(lat) / len(lat) * 1e-6),
                      "config": {"workload": "configs[3]: batch 1, 80 ms chunks (1280 samples), 16 kHz, 80 mel, 16 bins; "
                                             "chunk already on the device, latency = push() + stream sync; 5 frames per chunk",
                                 "note": "launch-latency bound: 5.5 KB per call, byte roofline not meaningful"}}))


def run_pool_workload(args, name):
    """configs[2] (calibrate + encode of 10k utterances, sharded) and configs[4] (long-form 2048)."""
    import torch.distributed as dist
    import dmel_codec_b200 as d
    from dmel_codec_b200 import distributed as D, synth
    rank, world, local_rank, dev = _init_rank()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    if name == "calibrate":
        geom = dict(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=256, n_mels=80)
        n_utts, seconds, n_bins, bsz, pool_n = 10000, 10, 16, 250, 500
        label = "configs[2]: min/max calibration + encode of 10,000 x 10 s utterances (16 kHz, 80 mel, 16 bins), block-sharded"
    else:
        geom = dict(sample_rate=44100, n_fft=2048, win_length=2048, hop_length=512, n_mels=160)
        n_utts, seconds, n_bins, bsz, pool_n = 32, 60, 32, 4, 4
        label = "configs[4]: long-form 32 x 60 s, 44.1 kHz, n_fft 2048, hop 512, 160 mel, 32 bins, utterances sharded"
    n = geom["sample_rate"] * seconds
    tok = d.DMelTokenizer(n_bins=n_bins, **geom).to(dev)
    mine = D.shard_range(n_utts, rank, world)
    # a pool of distinct synthetic utterances stands in for the shard (content does not change the work)
    pool = synth.device_batch(range(rank * pool_n, (rank + 1) * pool_n), n, geom["sample_rate"], dev)
    def load(ids):
        k = len(ids)
        start = (ids[0] * 7) % max(1, pool_n - k + 1) if k <= pool_n else 0
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
        dist.barrier()
    if rank == 0:
        audio_s = n_utts * seconds
        bytes_alg = 2 * 4 * n_utts * n + n_utts * geom["n_mels"] * (n // geom["hop_length"])
        peak, src = measured_peaks()
        print(json.dumps({"metric": f"dmel_{name}_encode_audio_seconds_per_second", "value": audio_s / (ms.item() / 1e3),
                          "unit": UNIT, "n_gpus": world, "ms_total": ms.item(), "higher_is_better": True,
                          "scaling": "strong", "hbm_frac_all_gpus": bytes_alg / (ms.item() / 1e3) / 1e9 / (peak * world),
                          "stats": {"lo_min": float(tok.quantizer.lo.min()), "hi_max": float(tok.quantizer.hi.max())},
                          "config": {"workload": label, "launch": tok._plan(dev).describe(),
                                     "job": "two transform passes" if two_pass else
                                            "one transform pass (log-mel of the shard kept in HBM) + stand-alone quantiser pass",
                                     "data": f"pool of {pool_n} distinct synthetic utterances per rank, cycled"}}))
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
    args = ap.parse_args()
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
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
