"""In-tree build of the CUDA shared libraries (sm_100a only; nvcc cross-compiles without a GPU)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]

TARGETS = {
    "libsrbtrack_b200.so": ["srb_api.cu", "fk_api.cu"],
}


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=True):
    nvcc = os.environ.get("NVCC", "nvcc")
    deps_all = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "srbtrack_b200.h"), os.path.join(HERE, "..", "include", "bonefk_b200.h")]
    built = []
    for lib, srcs in TARGETS.items():
        out = os.path.join(HERE, lib)
        if not force and not _stale(out, deps_all):
            continue
        cmd = [nvcc] + NVCC_FLAGS + ["-o", out] + [os.path.join(CSRC, s) for s in srcs]
        if verbose:
            print("[build]", " ".join(cmd), flush=True)
        subprocess.check_call(cmd)
        built.append(lib)
    return built


if __name__ == "__main__":
    build(force=True)
