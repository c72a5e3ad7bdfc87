"""bench.py's reference arm (`--impl reference`): real update() iterations of the CPU port, every host core, --steps and
--warmup honoured, one JSON line with the contract's keys.  Run here on the small C1-size workload (the driver runs it on
C4); started under OMP_NUM_THREADS=1 like a torchrun worker to prove that the arm undoes that."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_runs_real_iterations_on_all_cores():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "4",
                        "--warmup", "2", "--gpus", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "GiNGR update() iterations/s" and line["unit"] == "iterations/s"
    assert line["steps"] == 4 and line["warmup"] == 2                      # honoured, not re-interpreted
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert line["value"] > 0 and abs(line["ms_per_step"] * line["value"] - 1e3) < 1e-6 * 1e3
    assert line["e2e"] == {"value": line["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "no sampling, no extrapolation" in line["cpu_baseline"]["sample"] and line["gpu_launches"] == 0


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "2",
                        "--warmup", "1", "--gpus", "2"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""
