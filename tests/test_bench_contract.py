"""bench.py contract checks that need no GPU: the reference arm runs the reference's own file from baseline/_ref
(the oracle port where that is absent) and prints one JSON line with the agreed keys; both arms carry the same
`config`; the headline constants match BASELINE.json configs[1]."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["gpu_launches"] == 0
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["metric"] == "dmel_encode_audio_seconds_per_second" and line["unit"] == "audio-s/s"
    from baseline import ref_arm
    assert line["value"] > 0 and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["kind"] == ("reference" if ref_arm.available() else "port")
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.config_block()  # the key-for-key config our own arm prints
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]


def test_headline_workload_is_baseline_config_1():
    sys.path.insert(0, ROOT)
    import bench
    assert (bench.SAMPLE_RATE, bench.SECONDS, bench.BATCH, bench.N_BINS) == (24000, 10, 64, 16)
    assert bench.GEOM["n_fft"] == 1024 and bench.GEOM["hop_length"] == 256 and bench.GEOM["n_mels"] == 128
    assert bench.ENCODE_BYTES == 4 * 64 * 240000 + 64 * 128 * 937
    assert bench.DEQUANT_BYTES == 5 * 64 * 128 * 937
    assert bench.FORWARD_BYTES == bench.ENCODE_BYTES + 4 * 64 * 128 * 937   # codes are not re-read
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        assert "24 kHz speech, 128 mel, 16 bins, batch 64" in json.load(f)["configs"][1]
