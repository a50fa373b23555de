"""bench.py contract pieces that run without a GPU: the reference arm (`--impl reference`, the unmodified reference on host
cores) prints ONE JSON line with the keys the driver reads; our arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT, has_gpu


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=900, cwd=ROOT)


def test_reference_arm_line(ref):
    out = _run("--impl", "reference", "--workload", "rand_16MiB", "--steps", "1", "--warmup", "0")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sa_bwt_input_throughput" and d["unit"] == "MB/s"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"] == "rand_16MiB"
    # the reference arm runs on the WHOLE text of the configuration (same config as our arm) and also reports its inverse BWT
    assert d["config"]["sample_bytes"] == d["config"]["n_bytes"] == 1 << 24
    assert d["unbwt"]["value"] > 0 and d["unbwt"]["unit"] == "MB/s"


def test_default_mode_for_several_gpus_is_the_sharded_text():
    """`bench.py --gpus N` under torchrun must measure ONE text sharded over the ranks (strong scaling), not N replicas"""
    import re
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert re.search(r'add_argument\("--mode", default="sharded"', src)
    assert re.search(r'add_argument\("--isa", default="peer"', src)
    assert '"scaling": "strong"' in src


@pytest.mark.skipif(has_gpu(), reason="no-GPU behaviour")
def test_our_arm_fails_loudly_without_a_gpu():
    out = _run("--steps", "1", "--warmup", "0", "--n", "4096")
    assert out.returncode != 0
    assert "no CUDA device" in (out.stderr + out.stdout) or "no CPU fallback" in (out.stderr + out.stdout)
