"""world_size-2 gloo tests (CPU) of the N>1 host logic: utterance sharding and
the calibration all-reduce.  The per-shard statistics come from the oracle (the
CUDA kernels cannot run here); what is under test is that sharding covers every
utterance once and that MIN/MAX all-reduce of shard statistics is bit-identical
to single-process calibration."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN_GEOMETRY, oracle_config
from dmel_codec_b200 import distributed as D


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 10000, 10001):
        for w in (1, 2, 3, 8):
            got = [i for r in range(w) for i in D.shard_range(n, r, w)]
            assert got == list(range(n))
            sizes = [len(D.shard_range(n, r, w)) for r in range(w)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_range(4, 2, 2)
    assert [list(b) for b in D.batches(range(5), 2)] == [[0, 1], [2, 3], [4]]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, n_utts, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        import dmel_codec_b200 as d
        from dmel_codec_b200 import synth
        from oracle import dmel_oracle as O
        kw = GOLDEN_GEOMETRY["cfg1_16k_80"]
        cfg = oracle_config(kw)
        assert D.world() == (rank, world_size)
        mine = D.shard_range(n_utts, rank, world_size)
        q = d.DMelQuantizer(kw["n_mels"], 16)  # CPU buffers: gloo reduces them in place
        for ids in D.batches(mine, 3):
            lengths = torch.tensor([6000 + 500 * (i % 4) for i in ids])
            wav = synth.batch(ids, 8000, 16000, "speech", lengths=lengths.tolist())
            mel = O.log_mel(wav, cfg)
            lo, hi = O.calibrate_minmax(mel, O.valid_frames(lengths, cfg.hop_length))
            q.set_stats(torch.minimum(q.lo, lo), torch.maximum(q.hi, hi))
        q.sync_stats()
        torch.save({"lo": q.lo, "hi": q.hi, "n": len(mine)}, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_sharded_calibration_allreduce_is_bit_identical(tmp_path):
    from dmel_codec_b200 import synth
    from oracle import dmel_oracle as O
    n_utts, world_size = 10, 2
    mp.spawn(_worker, args=(world_size, _free_port(), n_utts, str(tmp_path)), nprocs=world_size, join=True)
    kw = GOLDEN_GEOMETRY["cfg1_16k_80"]
    cfg = oracle_config(kw)
    ids = list(range(n_utts))
    lengths = torch.tensor([6000 + 500 * (i % 4) for i in ids])
    mel = O.log_mel(synth.batch(ids, 8000, 16000, "speech", lengths=lengths.tolist()), cfg)
    lo, hi = O.calibrate_minmax(mel, O.valid_frames(lengths, cfg.hop_length))
    got = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world_size)]
    assert sum(g["n"] for g in got) == n_utts
    for g in got:
        assert torch.equal(g["lo"], lo) and torch.equal(g["hi"], hi)


class _OracleTokenizer:
    """CPU stand-in with the three members distributed.calibrate_encode_sharded uses; the transform and the
    quantiser are the oracle's (the CUDA kernels cannot run here).  Under test: the host logic of the
    single-transform job under world_size 2."""

    def __init__(self, kw, n_bins):
        import dmel_codec_b200 as d
        from oracle import dmel_oracle as O
        self.O, self.cfg, self.n_bins = O, oracle_config(kw), n_bins
        self.hop_length = kw["hop_length"]
        self.quantizer = d.DMelQuantizer(kw["n_mels"], n_bins)
        def encode(mel, mel_lengths=None, check_after=False):
            codes = O.dmel_encode(mel, self.quantizer.lo, self.quantizer.hi, n_bins)
            if mel_lengths is not None:
                codes = codes * (torch.arange(codes.shape[2])[None, None, :] < mel_lengths.reshape(-1)[:, None, None])
            return codes

        self.quantizer.encode = encode

    def n_frames(self, n_samples):
        return self.cfg.n_frames(n_samples)

    def update_stats_keep_mel(self, audios, lengths=None, out=None):
        mel = self.O.log_mel(audios, self.cfg)
        n_valid = None if lengths is None else self.O.valid_frames(lengths, self.cfg.hop_length)
        lo, hi = self.O.calibrate_minmax(mel, n_valid)
        self.quantizer.set_stats(torch.minimum(self.quantizer.lo, lo), torch.maximum(self.quantizer.hi, hi))
        if out is None:
            return mel
        assert out.shape == mel.shape and out.is_contiguous()
        out.copy_(mel)
        return out


def _lengths_of(ids):
    return torch.tensor([6000 + 500 * (i % 4) for i in ids])


def _padded(ids, ragged_tail):
    return 9000 if ragged_tail and min(ids) >= 8 else 8000


def _job_worker(rank, world_size, port, n_utts, out_dir, ragged_tail=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        from dmel_codec_b200 import synth
        tok = _OracleTokenizer(GOLDEN_GEOMETRY["cfg1_16k_80"], 16)

        def load(ids):
            lengths = _lengths_of(ids)
            return synth.batch(ids, _padded(ids, ragged_tail), 16000, "speech", lengths=lengths.tolist()), lengths

        out = [(list(ids), codes, code_lengths) for ids, codes, code_lengths in D.calibrate_encode_sharded(tok, n_utts, load, 3)]
        torch.save(out, os.path.join(out_dir, f"job{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ragged_tail", [False, True])
def test_single_transform_job_under_two_ranks(tmp_path, ragged_tail):
    """calibrate_encode_sharded on 2 ranks == tokenising the whole set in one process with dataset-wide
    statistics: every utterance once, in shard order, codes zero past the valid frames.  ragged_tail: the last batch
    of rank 1 is padded to another length, so it cannot share the shard's log-mel store and takes the per-batch path."""
    from dmel_codec_b200 import synth
    from oracle import dmel_oracle as O
    n_utts, world_size = 10, 2
    mp.spawn(_job_worker, args=(world_size, _free_port(), n_utts, str(tmp_path), ragged_tail), nprocs=world_size, join=True)
    kw = GOLDEN_GEOMETRY["cfg1_16k_80"]
    cfg = oracle_config(kw)
    ids = list(range(n_utts))
    lengths = _lengths_of(ids)
    mel = O.log_mel(synth.batch(ids, 9000, 16000, "speech", lengths=lengths.tolist()), cfg)
    n_valid = O.valid_frames(lengths, cfg.hop_length)
    lo, hi = O.calibrate_minmax(mel, n_valid)
    want = O.dmel_encode(mel, lo, hi, 16)
    want = want * (torch.arange(want.shape[2])[None, None, :] < n_valid[:, None, None])
    seen = []
    for r in range(world_size):
        for batch_ids, codes, code_lengths in torch.load(tmp_path / f"job{r}.pt"):
            t = cfg.n_frames(_padded(batch_ids, ragged_tail))
            assert codes.shape[2] == t
            assert torch.equal(codes, want[batch_ids][:, :, :t]) and torch.equal(code_lengths, n_valid[batch_ids])
            seen += batch_ids
    assert seen == ids
