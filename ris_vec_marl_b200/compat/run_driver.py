"""Run an UNMODIFIED reference driver script against the CUDA-backed `Environment` module.

    python -m ris_vec_marl_b200.compat.run_driver marl /path/to/reference [driver args...]
    python -m ris_vec_marl_b200.compat.run_driver sarl /path/to/reference
    python -m ris_vec_marl_b200.compat.run_driver marl_test /path/to/reference

The driver .py files and config.yaml are copied to a scratch directory (the reference
creates checkpoint directories next to its modules), the scratch dir
becomes the CWD, and sys.path is ordered: compat Environment > matplotlib stub > scratch.
"""
import os
import runpy
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DRIVERS = {"marl": ("Simulation-MARL-BCD", "marl_train_bcd.py"), "sarl": ("Simulation-SARL", "ddpg_train.py"),
           # the stale MADDPG test script: compat/marl also provides its missing `ddpg_torch` / `global_critic`
           "marl_test": ("Simulation-MARL-BCD", "marl_test.py")}


def main(argv):
    if len(argv) < 2 or argv[0] not in DRIVERS:
        raise SystemExit(__doc__)
    variant, ref_root, rest = argv[0], argv[1], argv[2:]
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
    sys.argv = [script] + rest
    runpy.run_path(os.path.join(scratch, script), run_name="__main__")


if __name__ == "__main__":
    main(sys.argv[1:])
