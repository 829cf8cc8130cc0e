"""`bench.py --impl reference` needs no GPU: run it at a small size and check the JSON line the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--ref-budget", "6", "--nside", "128", "--lmax", "383"] + extra,
                         capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_line():
    lines = _run([])
    assert len(lines) == 1
    d = lines[0]
    assert d["impl"] == "reference" and d["unit"] == "ms" and d["higher_is_better"] is False
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["value"] == d["ms_per_step"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["unit"] == "ms"
    # the thread count is what OpenMP grants, not what the machine has on paper
    assert 1 <= cb["cores"] <= (os.cpu_count() or 1)
    assert cb["sample_stride"] >= 1 and cb["extrapolated"] == (cb["sample_stride"] > 1)
    assert "Legendre" in cb["sample"] and "ring stage" in cb["sample"]


def test_reference_arm_under_torchrun_environment():
    """torchrun exports OMP_NUM_THREADS=1 and RANK/WORLD_SIZE: rank 0 still uses the machine, the other ranks print
    nothing and exit 0."""
    env = {"OMP_NUM_THREADS": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "0", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29999"}
    d = _run(["--gpus", "2"], dict(env, RANK="0"))
    assert len(d) == 1 and d[0]["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert _run(["--gpus", "2"], dict(env, RANK="1", LOCAL_RANK="1")) == []
