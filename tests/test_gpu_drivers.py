"""The three UNCHANGED reference driver scripts (marl_train_bcd.py, ddpg_train.py, the stale marl_test.py)
driving the CUDA-backed `Environment` module on a B200, compared with the same script, same seed, driving the
reference's own numpy `Environment` on the CPU (call sites marl_train_bcd.py:543-545,1611,
ddpg_train.py:40-42,162, marl_test.py:41,192-193).

Needs the reference tree: /root/reference (build container) or the git-ignored copy that
`tools/stage_reference.py` leaves under baseline/_ref, which travels to the GPU box.  Skipped LOUDLY without it.

What is compared per env step (ris_vec_marl_b200/compat/run_driver.py --record):
  * the state of the global numpy stream after the step -- EXACT on every step: the compat object consumes
    exactly the draws the reference consumes (arrivals, mobility, replay sampling of the learner);
  * DataBuf, data_t, data_p and the rewards -- within float32-vs-float64 tolerance while the two runs are
    still driven by the same actions (the learners are chaotic amplifiers once updates start, so values are
    compared over the first `compare_steps` steps; a QoS / penalty threshold flip moves a reward by a whole
    penalty, so a small fraction of reward samples may differ)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import ref_harness as rh

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOG_DIR = os.path.join(ROOT, "gpurun_out")


def run(variant, backend, record, max_steps, extra):
    cmd = [sys.executable, "-m", "ris_vec_marl_b200.compat.run_driver", variant, rh.REFERENCE_ROOT, "--backend", backend,
           "--record", record, "--max-steps", str(max_steps)] + extra
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT, env=env)
    assert res.returncode == 0, f"{variant}/{backend} failed:\n{res.stdout[-1500:]}\n{res.stderr[-3000:]}"
    return res.stdout


CASES = [
    # variant, env steps to run, steps whose values are compared, driver arguments
    ("marl", 230, 100, ["--", "--seed", "3", "--log", "none"]),   # 2 episodes + learner updates from step 128 on
    ("sarl", 230, 60, ["--seed", "5"]),                            # DDPG updates start at step 64
    ("marl_test", 150, 150, ["--seed", "7"]),                      # shipped MADDPG actors, inference only
]


@pytest.mark.parametrize("variant,steps,compare_steps,extra", CASES, ids=[c[0] for c in CASES])
def test_unchanged_driver_on_gpu_env_matches_reference_env(variant, steps, compare_steps, extra, tmp_path):
    if not rh.reference_available():
        pytest.skip("REFERENCE TREE NOT FOUND (neither /root/reference nor baseline/_ref): run tools/stage_reference.py "
                    "in the build container so the unchanged drivers can be exercised on the GPU box")
    recs = {}
    for backend in ("cuda", "reference"):
        path = str(tmp_path / f"{variant}_{backend}.npz")
        out = run(variant, backend, path, steps, extra)
        recs[backend] = dict(np.load(path))
        if backend == "cuda":
            os.makedirs(LOG_DIR, exist_ok=True)
            with open(os.path.join(LOG_DIR, f"driver_{variant}_cuda.log"), "w") as fh:
                fh.write(out[-20000:])
    g, r = recs["cuda"], recs["reference"]
    assert int(g["steps"]) == int(r["steps"]) == steps
    # the numpy stream: identical after every single env step
    assert list(g["stream"]) == list(r["stream"]), "the CUDA-backed env consumed different numpy draws"
    n = compare_steps
    worst = {}
    for k, atol in (("DataBuf", 2e-4), ("data_t", 2e-4), ("data_p", 2e-4)):
        err = np.abs(g[k][:n] - r[k][:n])
        tol = atol + 1e-4 * np.abs(r[k][:n])
        frac_bad = float((err > tol).mean())
        worst[k] = float(err.max())
        assert frac_bad <= 0.02, f"{variant}: {k} differs on {frac_bad:.1%} of the samples (max {err.max():.3g})"
    rew_err = np.abs(g["reward"][:n] - r["reward"][:n])
    frac_bad = float((rew_err > 1e-3 + 1e-3 * np.abs(r["reward"][:n])).mean())
    assert frac_bad <= 0.05, f"{variant}: reward differs on {frac_bad:.1%} of the first {n} steps"
    line = (f"[drivers] {variant}: {steps} env steps through the unchanged script on both backends; numpy stream identical "
            f"after every step; first {n} steps: max |dDataBuf| {worst['DataBuf']:.2e}, max |ddata_t| {worst['data_t']:.2e}, "
            f"reward off on {frac_bad:.1%} of steps (median |d| {np.median(rew_err):.2e})")
    print(line)
    with open(os.path.join(LOG_DIR, "drivers_summary.log"), "a") as fh:
        fh.write(line + "\n")
