"""Builds liblaminr_b200.so (the C-ABI library of include/laminr_b200.h) in-tree with nvcc for sm_100a."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SOURCES = ["api.cu", "inr_simt.cu", "inr_tc.cu", "post.cu", "learn_obj.cu", "convcone.cu", "fullframe.cu", "mask.cu"]
OBJ_DIR = os.path.join(HERE, "_obj")  # git-ignored object cache: only changed translation units are recompiled
OUT = os.path.join(os.path.dirname(HERE), "liblaminr_b200.so")


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(ROOT, "include", "laminr_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def _stale(obj, src, headers):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in [src] + headers)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    from concurrent.futures import ThreadPoolExecutor

    nvcc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-Xcompiler", "-fPIC"]
    if verbose:
        flags.append("-Xptxas=-v")
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".cuh")] + [os.path.join(ROOT, "include", "laminr_b200.h")]
    objs, jobs = [], []
    for s in SOURCES:
        src, obj = os.path.join(HERE, s), os.path.join(OBJ_DIR, s[:-3] + ".o")
        objs.append(obj)
        if force or verbose or _stale(obj, src, headers):
            jobs.append([nvcc] + flags + ["-c", src, "-o", obj])
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        for r in ex.map(lambda c: subprocess.run(c), jobs):
            if r.returncode != 0:
                raise subprocess.CalledProcessError(r.returncode, r.args)
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", OUT] + objs, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
