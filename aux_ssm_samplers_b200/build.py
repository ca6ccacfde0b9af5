"""Builds libauxssm.so in-tree with nvcc for sm_100a (no torch, no JIT cache).

`python -m aux_ssm_samplers_b200.build` or `build_library()`; object files are compiled in
parallel (one nvcc process per .cu) and only when their sources changed.
"""
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "libauxssm.so")
OBJ_DIR = os.path.join(CSRC, "_obj")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp(paths):
    h = hashlib.sha1()
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _compile_one(nvcc, src, headers):
    obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
    stamp_file = obj + ".stamp"
    deps = [os.path.join(CSRC, src)] + headers
    if src.endswith("_f64.cu"):  # X_f64.cu is "#define AUXSSM_ONLY_F64 / #include X.cu"
        deps.append(os.path.join(CSRC, src[:-7] + ".cu"))
    stamp = _stamp(deps)
    if os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return obj, False
    cmd = [nvcc] + NVCC_FLAGS + ["-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return obj, True


def build_library(verbose=False, force=False):
    """Compile every csrc/*.cu for sm_100a and link libauxssm.so next to this file."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    headers = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh"))
    headers.append(os.path.join(INCLUDE, "auxssm.h"))
    if force:
        for f in os.listdir(OBJ_DIR):
            os.remove(os.path.join(OBJ_DIR, f))
    srcs = _sources()
    # the slowest translation units first, so that they overlap with everything else on the worker pool
    heavy = ["kalman_filter.cu", "kalman_filter_f64.cu", "kalman_sample.cu", "kalman_sample_f64.cu", "lgssm_logpdf.cu",
             "lgssm_logpdf_f64.cu", "csmc_guided.cu", "csmc_independent.cu", "csmc_bootstrap.cu"]
    srcs = [h for h in heavy if h in srcs] + [x for x in srcs if x not in heavy]
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile_one(nvcc, s, headers), srcs))
    objs = [o for o, _ in results]
    rebuilt = any(r for _, r in results)
    if rebuilt or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        print("libauxssm.so:", "rebuilt" if rebuilt else "up to date", LIB)
    return LIB


def build_xla_ffi(verbose=False):
    """Builds libauxssm_xla.so (csrc/xla/xla_ffi.cc: one XLA FFI handler per entry point) next to libauxssm.so.
    Needs jaxlib's headers (``jax.ffi.include_dir()``); returns None where JAX is not installed (this image)."""
    try:
        import jax.ffi
        inc = jax.ffi.include_dir()
    except Exception as e:  # noqa: BLE001
        if verbose:
            print("xla_ffi adapter skipped: jax.ffi is not importable here (%s)" % e)
        return None
    out = os.path.join(HERE, "libauxssm_xla.so")
    cmd = ["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-I", inc, "-I", INCLUDE, "-I", "/usr/local/cuda/include",
           os.path.join(CSRC, "xla", "xla_ffi.cc"), "-o", out, "-L", HERE, "-lauxssm", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("xla_ffi build failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        print("libauxssm_xla.so built:", out)
    return out


if __name__ == "__main__":
    build_library(verbose=True, force="--force" in sys.argv)
    if "--xla-ffi" in sys.argv:
        build_xla_ffi(verbose=True)
