"""In-tree build of the CUDA shared libraries (sm_100a only; nvcc cross-compiles without a GPU).

Every `.cu` is compiled to its own object (in parallel, only when it or a header changed) and the objects are linked
into one shared library, so touching one kernel file does not rebuild the others."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
INCLUDE = os.path.join(HERE, "..", "include")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]

TARGETS = {
    "libsrbtrack_b200.so": ["srb_api.cu", "fk_api.cu", "cdm_api.cu", "mmsto_api.cu"],
}
# headers each translation unit depends on (besides itself)
DEPS = {
    "srb_api.cu": ["srb_kernels.cuh", "srb_math.cuh", "srb_types.h", "../../include/srbtrack_b200.h"],
    "fk_api.cu": ["fk_kernels.cuh", "srb_math.cuh", "../../include/bonefk_b200.h"],
    "mmsto_api.cu": ["mmsto_kernels.cuh", "../../include/mmsto_b200.h"],
    "cdm_api.cu": ["cdm_kernels.cuh", "cdm_types.h", "srb_kernels.cuh", "srb_math.cuh", "srb_types.h", "../../include/cdmenv_b200.h"],
}


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=True):
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    built = []
    for lib, srcs in TARGETS.items():
        procs = []
        objs = []
        for s in srcs:
            src = os.path.join(CSRC, s)
            obj = os.path.join(OBJ, s.replace(".cu", ".o"))
            objs.append(obj)
            deps = [src] + [os.path.normpath(os.path.join(CSRC, d)) for d in DEPS.get(s, [])]
            if force or _stale(obj, deps):
                cmd = [nvcc] + NVCC_FLAGS + ["-c", "-o", obj, src]
                if verbose:
                    print("[build]", " ".join(cmd), flush=True)
                procs.append((cmd, subprocess.Popen(cmd)))
        failed = [cmd for cmd, p in procs if p.wait() != 0]
        if failed:
            raise subprocess.CalledProcessError(1, failed[0])
        out = os.path.join(HERE, lib)
        if force or procs or _stale(out, objs):
            cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs
            if verbose:
                print("[build]", " ".join(cmd), flush=True)
            subprocess.check_call(cmd)
            built.append(lib)
    return built


if __name__ == "__main__":
    build(force=True)
