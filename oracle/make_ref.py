"""make_ref.py — TEST / BENCH INFRASTRUCTURE.  Stages the UNMODIFIED reference package for the GPU box.

The reference (Midren/rl_sandbox) is pure Python: there is nothing to compile.  What `--impl reference` and the
`torch_gpu_baseline` leg of bench.py need on the GPU box — where /root/reference does not exist — is the package itself,
so this recipe copies the .py files of the hot path's import closure (rl_sandbox/{agents,utils,vision}) verbatim from
/root/reference into oracle/_ref/rl_sandbox/.  oracle/_ref/ is git-ignored (the sources never enter the history) but not
gpurun-ignored, so it travels like the built .so files.  Nothing in the product path imports it.

Run:  python -m oracle.make_ref        (also run by __graft_entry__.build() when /root/reference is present)
"""
import shutil
from pathlib import Path

SRC = Path("/root/reference/rl_sandbox")
DST = Path(__file__).resolve().parent / "_ref" / "rl_sandbox"
PARTS = ["__init__.py", "agents", "utils", "vision"]


def main() -> bool:
    if not SRC.exists():
        return False
    if DST.exists():
        shutil.rmtree(DST)
    DST.mkdir(parents=True)
    for part in PARTS:
        s = SRC / part
        if s.is_dir():
            shutil.copytree(s, DST / part, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        elif s.exists():
            shutil.copy2(s, DST / part)
    (DST.parent / "README").write_text(
        "Verbatim copy of /root/reference/rl_sandbox/{agents,utils,vision} made by oracle/make_ref.py (git-ignored).\n")
    return True


if __name__ == "__main__":
    print("oracle/_ref staged" if main() else "no /root/reference here: nothing staged")
