"""Run an UNMODIFIED reference driver script against the CUDA-backed `Environment` module.

    python -m ris_vec_marl_b200.compat.run_driver marl /path/to/reference [options] [-- driver args...]
    python -m ris_vec_marl_b200.compat.run_driver sarl /path/to/reference
    python -m ris_vec_marl_b200.compat.run_driver marl_test /path/to/reference

The driver .py files and config.yaml are copied to a scratch directory (the reference
creates checkpoint directories next to its modules), the scratch dir
becomes the CWD, and sys.path is ordered: compat Environment > matplotlib stub > scratch.

Options (for tests/test_gpu_drivers.py; none of them touches a driver file):
    --backend cuda|reference   which `Environment` module the script imports: the CUDA-backed compat one
                               (default) or the reference's own numpy module (the CPU run to compare with;
                               for the stale one-argument `marl_test.py` its `step(action)` is completed
                               with all-singleton groups exactly as the compat object does)
    --record FILE.npz          log every `Environ.step` return value and the position of the global numpy
                               stream after it
    --max-steps N              stop the driver cleanly after N env steps
    --seed S                   seed numpy / random / torch right before the script starts (scripts that
                               do not seed themselves: ddpg_train.py, marl_test.py)
"""
import hashlib
import importlib
import importlib.util
import os
import random
import runpy
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DRIVERS = {"marl": ("Simulation-MARL-BCD", "marl_train_bcd.py"), "sarl": ("Simulation-SARL", "ddpg_train.py"),
           # the stale MADDPG test script: compat/marl also provides its missing `ddpg_torch` / `global_critic`
           "marl_test": ("Simulation-MARL-BCD", "marl_test.py")}


class StopDriver(SystemExit):
    """Raised from inside `Environ.step` once --max-steps env steps have run."""


class Recorder:
    """Wraps `Environ.step` of the imported `Environment` module (either backend)."""

    def __init__(self, mod, variant, path, max_steps):
        import numpy as np

        self.np, self.path, self.max_steps, self.variant = np, path, max_steps, variant
        self.rows = {"reward": [], "DataBuf": [], "data_t": [], "data_p": [], "reward_user": [], "stream": []}
        self.n = 0
        inner = mod.Environ.step
        rec = self

        def step(env, action_power, second=None):
            if variant == "sarl":
                out = inner(env, action_power, second)
                reward, buf, dt, dp = out[0], out[1], out[2], out[3]
                ru = np.zeros(len(buf))
            else:
                if second is None:  # marl_test.py:192 calls step(action): every user its own group
                    second = [[i] for i in range(env.n_veh)]
                out = inner(env, action_power, second)
                ru, reward, buf, dt, dp = out[0], out[1], out[2], out[3], out[4]
            st = np.random.get_state()
            digest = hashlib.sha1(st[1].tobytes()).hexdigest()[:16] + f":{st[2]}"
            r = rec.rows
            r["reward"].append(float(reward)); r["DataBuf"].append(np.array(buf, dtype=float))
            r["data_t"].append(np.array(dt, dtype=float)); r["data_p"].append(np.array(dp, dtype=float))
            r["reward_user"].append(np.array(ru, dtype=float)); r["stream"].append(digest)
            rec.n += 1
            if rec.max_steps and rec.n >= rec.max_steps:
                raise StopDriver(0)
            return out

        mod.Environ.step = step

    def save(self):
        if self.path:
            np = self.np
            self.np.savez(self.path, **{k: np.array(v) for k, v in self.rows.items()}, steps=self.n)


def main(argv):
    if len(argv) < 2 or argv[0] not in DRIVERS:
        raise SystemExit(__doc__)
    variant, ref_root, rest = argv[0], argv[1], argv[2:]
    opts = {"--backend": "cuda", "--record": None, "--max-steps": "0", "--seed": None}
    while rest and rest[0] in opts:
        opts[rest[0]] = rest[1]
        rest = rest[2:]
    if rest and rest[0] == "--":
        rest = rest[1:]
    backend = opts["--backend"]
    sub, script = DRIVERS[variant]
    src = os.path.join(ref_root, sub)
    scratch = tempfile.mkdtemp(prefix=f"risvec_{variant}_")
    for f in os.listdir(src):
        if f.endswith((".py", ".yaml")) and f != "Environment.py":
            shutil.copy(os.path.join(src, f), scratch)
    os.chdir(scratch)
    compat_dir = "marl" if variant == "marl_test" else variant
    if variant == "marl_test":  # shipped MADDPG checkpoints stay where they are (read-only)
        os.environ.setdefault("RISVEC_MARL_MODEL_DIR", os.path.join(src, "model2", "3-BCD_RIS_marl_ddpg-8"))
    sys.path[:0] = [os.path.join(HERE, compat_dir), os.path.join(HERE, "stubs"), scratch]
    if backend == "reference":  # the reference's own numpy module, loaded from where it lies
        spec = importlib.util.spec_from_file_location("Environment", os.path.join(src, "Environment.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["Environment"] = mod
        dont = sys.dont_write_bytecode
        sys.dont_write_bytecode = True
        try:
            spec.loader.exec_module(mod)
        finally:
            sys.dont_write_bytecode = dont
    else:
        mod = importlib.import_module("Environment")
    recorder = None
    if opts["--record"] or int(opts["--max-steps"]) or variant == "marl_test" and backend == "reference":
        recorder = Recorder(mod, "sarl" if variant == "sarl" else "marl", opts["--record"], int(opts["--max-steps"]))
    if opts["--seed"] is not None:
        import numpy as np
        import torch

        seed = int(opts["--seed"])
        np.random.seed(seed); random.seed(seed); torch.manual_seed(seed)
    sys.argv = [script] + rest
    try:
        runpy.run_path(os.path.join(scratch, script), run_name="__main__")
    except StopDriver:
        print(f"[run_driver] stopped after {recorder.n} env steps (--max-steps)")
    finally:
        if recorder is not None:
            recorder.save()


if __name__ == "__main__":
    main(sys.argv[1:])
