"""In-tree nvcc build of the CUDA libraries (sm_100a only; no JIT cache, the .so travels with the tree)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

LIBS = {
    # output (relative to the package dir) : sources (relative to csrc/)
    "librp_b200.so": ["rp_solver.cu", "scnet.cu", "scnet_tc.cu", "scnet_halo.cu", "rp_warp.cu", "rp_keypoint.cu", "rp_plan.cu"],
}


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(out, srcs):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    deps = list(srcs) + [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    return any(os.path.getmtime(d) > t for d in deps)


def build_all(force=False, verbose=False):
    built = []
    for out, srcs in LIBS.items():
        outp = os.path.join(HERE, out)
        srcp = [os.path.join(CSRC, s) for s in srcs]
        if not force and not _stale(outp, srcp):
            continue
        cmd = [_nvcc()] + NVCC_FLAGS + ["-I", INCLUDE, "-o", outp] + srcp
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
            print(" ".join(cmd))
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (out, r.stdout))
        if verbose:
            print(r.stdout)
        built.append(out)
    return built


if __name__ == "__main__":
    print("built:", build_all(force="--force" in sys.argv, verbose=True))
