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
    "librp_b200.so": ["rp_solver.cu", "rp_solver_wide.cu", "rp_host.cu", "scnet.cu", "scnet_tc.cu", "scnet_halo.cu", "rp_warp.cu", "rp_keypoint.cu", "rp_plan.cu"],
}


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build_host_ext(force=False):
    """The CPython extension _rp_pack (csrc/rp_pack.c: host-side packing of record lists; pure C, no CUDA) with gcc, in-tree."""
    import sysconfig
    src = os.path.join(CSRC, "rp_pack.c")
    out = os.path.join(HERE, "_rp_pack.so")
    if not (force or _stale(out, [src])):
        return []
    cc = os.environ.get("CC") or shutil.which("gcc") or shutil.which("cc")
    if not cc:
        raise RuntimeError("no C compiler for _rp_pack")
    cmd = [cc, "-O2", "-fPIC", "-shared", "-Wall", "-I", sysconfig.get_paths()["include"], "-o", out, src]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("building _rp_pack failed:\n%s" % r.stdout)
    return ["_rp_pack.so"]


def build_all(force=False, verbose=False):
    """Each .cu is compiled to an object (only when stale, in parallel), then linked into the one shared library."""
    from concurrent.futures import ThreadPoolExecutor
    built = []
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    headers += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    for out, srcs in LIBS.items():
        outp = os.path.join(HERE, out)
        jobs, objs = [], []
        for sname in srcs:
            src = os.path.join(CSRC, sname)
            obj = os.path.join(objdir, sname[:-3] + ".o")
            objs.append(obj)
            inc = [os.path.join(CSRC, ln.split('"')[1]) for ln in open(src) if ln.startswith('#include "') and ln.split('"')[1].endswith(".cu")]
            if force or _stale(obj, [src] + inc + headers):
                cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f != "-shared"] + os.environ.get("RP_NVCC_EXTRA", "").split() + ["-I", INCLUDE, "-c", "-o", obj, src]
                if verbose:
                    cmd[1:1] = ["-Xptxas", "-v"]
                jobs.append((sname, cmd))

        def run(job):
            r = subprocess.run(job[1], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            return job[0], r.returncode, r.stdout
        with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
            for name, rc, log in ex.map(run, jobs):
                if rc != 0:
                    raise RuntimeError("nvcc failed for %s:\n%s" % (name, log))
                if verbose:
                    print("== %s\n%s" % (name, log))
        if jobs or force or _stale(outp, objs):
            r = subprocess.run([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", outp] + objs,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if r.returncode != 0:
                raise RuntimeError("link failed for %s:\n%s" % (out, r.stdout))
            built.append(out)
    built += build_host_ext(force)
    return built


if __name__ == "__main__":
    print("built:", build_all(force="--force" in sys.argv, verbose=True))
