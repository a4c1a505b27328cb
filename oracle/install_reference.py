"""TEST / BENCH INFRASTRUCTURE ONLY -- the one offline install of the reference (bench contract: "reference arm").

The reference tree has neither setup.py nor pyproject.toml, so ``pip install --target baseline/_ref /root/reference``
fails ("not installable").  Following the contract ("if the build needs to write into the source tree, install from a
copy under /tmp"), this script copies the packages the hot path imports (RPModule/, utils/, model/, util.py, config.py -- nothing
else) to a scratch directory, adds a minimal setup.py there, and runs the prescribed

    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref <copy>

``baseline/_ref/`` is git-ignored (the reference's sources never enter this repository's history) but not
gpurun-ignored, so the UNMODIFIED files travel to the GPU box, where ``bench.py --impl reference`` and ``cpu_baseline``
execute them through oracle/ref_loader.py (in-memory syntax shim for rpmodule.py:342-343 only).  Idempotent; a no-op
when /root/reference is absent (the GPU box uses what was installed here).
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")

SETUP = """from setuptools import setup
setup(name="relativepose-reference", version="0", packages=["RPModule", "utils", "model"], py_modules=["util", "config"])
"""


def install(force=False):
    marker = os.path.join(DST, "RPModule", "rpmodule.py")
    if os.path.isfile(marker) and not force:
        return "present"
    if not os.path.isdir(SRC):
        return "no reference tree"
    tmp = tempfile.mkdtemp(prefix="rp_ref_src_")
    try:
        for name in ("RPModule", "utils", "model"):
            shutil.copytree(os.path.join(SRC, name), os.path.join(tmp, name),
                            ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        for mod in ("util.py", "config.py"):
            shutil.copy(os.path.join(SRC, mod), os.path.join(tmp, mod))
        with open(os.path.join(tmp, "setup.py"), "w") as fh:
            fh.write(SETUP)
        os.makedirs(DST, exist_ok=True)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--no-compile", "--upgrade",
               "--find-links", "/opt/wheelhouse", "--target", DST, tmp]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0 or not os.path.isfile(marker):
            return "pip install failed: " + r.stdout.strip().splitlines()[-1]
        return "installed"
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
