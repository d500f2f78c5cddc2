"""Multi-GPU plumbing of the dMel path: one process per GPU, utterances sharded
across ranks, and exactly one collective — the min/max all-reduce that turns
per-shard calibration statistics into dataset-wide bin edges.

Encode itself needs no communication: frames and utterances are independent.
The reference has no direct collectives either (Lightning DDP only, reference
config/codec/dMel_used.yaml:18); its per-rank data split is by the lhotse
sampler (reference dataset/lhotse_tts_dataset.py:184-191).
"""
from __future__ import annotations

from typing import Callable, Iterable, Iterator, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_items: int, rank: int, world_size: int) -> range:
    """Contiguous block of item ids owned by ``rank``; blocks differ by at most
    one item and cover ``range(n_items)`` exactly once."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def batches(ids: Sequence[int], batch_size: int) -> Iterator[Sequence[int]]:
    for i in range(0, len(ids), batch_size):
        yield ids[i:i + batch_size]


@torch.no_grad()
def calibrate_sharded(tokenizer, n_utterances: int, load_batch: Callable[[Sequence[int]], object],
                      batch_size: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Dataset-wide calibration: every rank scans its block of utterance ids
    (``load_batch(ids)`` returns audios or (audios, audio_lengths) on this
    rank's GPU), then lo/hi are all-reduced (MIN / MAX).  Exact and order
    independent, so the result is bit-identical for any number of ranks."""
    rank, size = world()
    mine = shard_range(n_utterances, rank, size)
    tokenizer.calibrate((load_batch(ids) for ids in batches(mine, batch_size)), group=group)
    return tokenizer.quantizer.lo, tokenizer.quantizer.hi


@torch.no_grad()
def encode_sharded(tokenizer, n_utterances: int, load_batch: Callable[[Sequence[int]], object],
                   batch_size: int) -> Iterable[Tuple[Sequence[int], torch.Tensor, Optional[torch.Tensor]]]:
    """Yield (utterance ids, codes, code_lengths) for this rank's block. No communication."""
    rank, size = world()
    for ids in batches(shard_range(n_utterances, rank, size), batch_size):
        item = load_batch(ids)
        audios, lengths = item if isinstance(item, (tuple, list)) else (item, None)
        codes, code_lengths = tokenizer.encode(audios, lengths)
        yield ids, codes, code_lengths


@torch.no_grad()
def calibrate_encode_sharded(tokenizer, n_utterances: int, load_batch: Callable[[Sequence[int]], object],
                             batch_size: int, group=None
                             ) -> Iterable[Tuple[Sequence[int], torch.Tensor, Optional[torch.Tensor]]]:
    """``calibrate_sharded`` followed by ``encode_sharded`` with the STFT done once: pass 1 writes the
    log-mel of this rank's block to HBM while it folds the statistics (one launch per batch), the
    statistics are all-reduced, and pass 2 is the stand-alone quantiser over the stored log-mel —
    an HBM-bound pass of 5 bytes per value instead of a second transform.  Same codes, bit for bit, as
    the two-pass job (the fused encode quantises exactly these float32 values).  Needs 4 bytes per
    log-mel value of the shard in HBM (2 GB for 10,000 x 10 s at 80 mel); when that is more than
    ``KEEP_MEL_HBM_FRACTION`` of the free HBM the job runs as the two transform passes instead."""
    rank, size = world()
    mine = shard_range(n_utterances, rank, size)
    def two_passes():  # the shard's log-mel does not fit next to the waveforms: the transform runs twice, same codes
        calibrate_sharded(tokenizer, n_utterances, load_batch, batch_size, group=group)
        yield from encode_sharded(tokenizer, n_utterances, load_batch, batch_size)

    q = tokenizer.quantizer
    q.reset_stats()
    # Pass 1.  Batches of the shard's common padded length write straight into ONE store, so that pass 2 is a single
    # launch over it (forty 15-microsecond launches are bound by the host, not by HBM); a batch of another length
    # keeps its own tensor and its own launch.
    store, used, kept = None, 0, []   # kept: (ids, lengths, row offset in the store | None, own log-mel | None)
    for ids in batches(mine, batch_size):
        item = load_batch(ids)
        audios, lengths = item if isinstance(item, (tuple, list)) else (item, None)
        b, t = (audios.shape[0] if audios.ndim > 1 else 1), tokenizer.n_frames(audios.shape[-1])
        if store is None:
            # the first batch gives the shard's frame count: decide here whether the store fits (every rank ends up with
            # exactly one all-reduce whichever way it decides, so ranks need not agree)
            try:
                if not _store_fits(4 * len(mine) * q.n_mels * t, audios.device):
                    raise torch.cuda.OutOfMemoryError("the shard's log-mel does not fit")
                store = _alloc_store((len(mine), q.n_mels, t), audios.device)
            except torch.cuda.OutOfMemoryError:
                del item, audios, lengths
                yield from two_passes()
                return
        if store is not None and t == store.shape[2] and used + b <= store.shape[0] and audios.device == store.device:
            tokenizer.update_stats_keep_mel(audios, lengths, out=store[used:used + b])
            kept.append((ids, lengths, used, None))
            used += b
        else:
            kept.append((ids, lengths, None, tokenizer.update_stats_keep_mel(audios, lengths)))
    q.sync_stats(group)
    # Pass 2: the stand-alone quantiser over the stored log-mel; frames past a valid length get code 0 unread
    codes_all = None
    if used:
        in_store = [k for k in kept if k[2] is not None]
        lens = None
        if all(k[1] is not None for k in in_store):
            lens = torch.div(torch.cat([k[1].reshape(-1).to(store.device) for k in in_store]), tokenizer.hop_length,
                             rounding_mode="floor")
        rows_per_call = max(1, ((1 << 31) - 1) // (store.shape[1] * store.shape[2]))  # the C entry indexes with 32 bits
        parts = [q.encode(store[r:min(r + rows_per_call, used)], None if lens is None else lens[r:min(r + rows_per_call, used)],
                          check_after=True) for r in range(0, used, rows_per_call)]
        codes_all = parts[0] if len(parts) == 1 else torch.cat(parts)
    for ids, lengths, at, own in kept:
        code_lengths = None
        if lengths is not None:
            code_lengths = torch.div(lengths.reshape(-1).to(store.device if at is not None else own.device), tokenizer.hop_length,
                                     rounding_mode="floor")
        if at is not None:
            codes = codes_all[at:at + len(ids)]
            if lens is None and code_lengths is not None:   # a store that mixes batches with and without lengths
                codes = _mask_past(codes, code_lengths)
        else:
            codes = q.encode(own, code_lengths)
        yield ids, codes, code_lengths


def _mask_past(codes: torch.Tensor, code_lengths: torch.Tensor) -> torch.Tensor:
    """codes with 0 at frames >= code_lengths[b]: what the fused encode writes past the valid frames."""
    t = torch.arange(codes.shape[2], device=codes.device)
    return codes * (t[None, None, :] < code_lengths[:, None, None])


# fraction of the free HBM the stored log-mel of a shard may take (the rest: waveform batches, codes, allocator slack)
KEEP_MEL_HBM_FRACTION = 0.6


def _alloc_store(shape, device) -> torch.Tensor:
    return torch.empty(shape, dtype=torch.float32, device=device)


_TOTAL_HBM = {}  # device index -> bytes (asked once: even the allocator's statistics cost 50 us per query)


def _store_fits(need: int, device) -> bool:
    """Whether ``need`` bytes of log-mel fit in this GPU's HBM next to everything else (DMEL_KEEP_MEL=0/1 overrides the
    decision).  The driver's free-memory query is a synchronising call that now and then takes milliseconds, so only
    the tight cases ask it: a store under a quarter of ``KEEP_MEL_HBM_FRACTION`` of the GPU's memory is taken to fit,
    and if the allocation then fails (this or another process holds the memory) the caller falls back."""
    import os
    forced = os.environ.get("DMEL_KEEP_MEL")
    if forced is not None:
        return forced != "0"
    device = torch.device(device)
    if device.type != "cuda":
        return True
    index = device.index if device.index is not None else torch.cuda.current_device()
    if index not in _TOTAL_HBM:
        _TOTAL_HBM[index] = torch.cuda.get_device_properties(index).total_memory
    if KEEP_MEL_HBM_FRACTION > 0 and need <= 0.25 * KEEP_MEL_HBM_FRACTION * _TOTAL_HBM[index]:
        return True
    free, _total = torch.cuda.mem_get_info(device)
    return need <= KEEP_MEL_HBM_FRACTION * free
