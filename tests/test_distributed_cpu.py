"""world_size-2 (and 4) CPU tests of the N > 1 host logic, gloo backend, launched exactly as the driver launches
bench.py (python -m torch.distributed.run, 127.0.0.1).  No GPU, no rendering: the data path has no collective."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def torchrun(nproc, port, *args, timeout=300):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--gpus", str(nproc), *args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout      # rank 0 alone prints
    return lines[0]


@pytest.mark.parametrize("nproc,port", [(2, 29631), (4, 29632)])
def test_sharding_and_reductions_over_gloo(nproc, port):
    j = torchrun(nproc, port, "--dry-run", "--workload", "spheres4k")
    assert j["dry_run"] and j["n_gpus"] == nproc and j["max_rank_plus_1"] == nproc
    assert j["samples_per_step"] == 3840 * 2160 * 1024          # the shards partition the frame exactly
    assert j["tiles"] == 60 * 68
    assert abs(j["my_tiles"] - j["tiles"] / nproc) <= 68        # balanced static interleave
    one = json.loads(subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--dry-run", "--workload", "spheres4k"],
                                    capture_output=True, text=True, cwd=ROOT).stdout.strip().splitlines()[-1])
    assert one["samples_per_step"] == j["samples_per_step"] and one["my_tiles"] == j["tiles"]


def test_reference_arm_under_torchrun_runs_on_rank0_only(oracle):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built")
    j = torchrun(2, 29633, "--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "spheres", "--ref-stride", "97", "--ref-spp", "4")
    assert j["impl"] == "reference" and j["value"] > 0 and j["unit"] == "Msamples/s" and j["metric"] == "Msamples/s"
    assert j["cpu_baseline"]["kind"] == "reference" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0


def test_committed_bench_lines_follow_the_contract():
    """The bench lines committed under profiles/ (produced on a B200 by `python bench.py` and `--impl reference`) carry
    every key the measurement contract names."""
    ours = json.loads(open(os.path.join(ROOT, "profiles", "r01_bench_default_4k_final.jsonl")).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in ours, k
    assert ours["metric"] == "Msamples/s" and ours["unit"] == "Msamples/s" and ours["dtype"] == "f32" and ours["vs_baseline"] is None
    assert "workload" in ours["config"] and "model" not in ours["config"]
    assert set(ours["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and ours["e2e"]["d2h_bytes_per_step"] > 0
    assert set(ours["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert set(ours["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and ours["cpu_baseline"]["kind"] == "reference"
    assert set(ours["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert ours["gpu_launches"] > 0 and ours["value"] > 1000
    ref = json.loads(open(os.path.join(ROOT, "profiles", "r01_bench_reference_arm.jsonl")).read().strip().splitlines()[-1])
    assert ref["impl"] == "reference" and ref["metric"] == ours["metric"] and ref["unit"] == ours["unit"]
    assert ref["config"]["workload"] == ours["config"]["workload"] and ref["higher_is_better"] == ours["higher_is_better"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["cpu_baseline"]["kind"] == "reference"
