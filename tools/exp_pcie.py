"""Experiment: D2H transfer primitives of one e2e step's outputs (4096 x 150 floats) on this box -- contiguous copy, pitched copy of
the 125-float window part, pitched copy of the 25-float self part, small arrays -- alone and while a compute kernel runs.
   python tools/exp_pcie.py"""
import ctypes as C, time, sys, os
import numpy as np, torch
rt = None
for name in ("libcudart.so.12", "libcudart.so"):
    try:
        rt = C.CDLL(name); break
    except OSError:
        pass
if rt is None:
    import glob
    rt = C.CDLL(glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*"))[0])
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
N, W = 4096, 150
d = torch.randn(N, W, device="cuda")
h = torch.empty(N, W).pin_memory()
ha = torch.randn(N, 22).pin_memory(); da = torch.empty(N, 22, device="cuda")
small_d = [torch.randn(N, device="cuda"), torch.zeros(N, dtype=torch.uint8, device="cuda"), torch.randn(N, 8, device="cuda")]
small_h = [torch.empty(N).pin_memory(), torch.empty(N, dtype=torch.uint8).pin_memory(), torch.empty(N, 8).pin_memory()]
s = torch.cuda.Stream(); sp = C.c_void_p(s.cuda_stream)
s2 = torch.cuda.Stream()
D2H, H2D = 2, 1
big = torch.randn(8192, 8192, device="cuda")
def busy():
    with torch.cuda.stream(s2):
        for _ in range(2): torch.mm(big, big)
def full(): rt.cudaMemcpyAsync(h.data_ptr(), d.data_ptr(), N * W * 4, D2H, sp)
def window(): rt.cudaMemcpy2DAsync(h.data_ptr() + 100, W * 4, d.data_ptr() + 100, W * 4, 500, N, D2H, sp)
def selfp(): rt.cudaMemcpy2DAsync(h.data_ptr(), W * 4, d.data_ptr(), W * 4, 100, N, D2H, sp)
def small():
    for a, b in zip(small_h, small_d): rt.cudaMemcpyAsync(a.data_ptr(), b.data_ptr(), b.numel() * b.element_size(), D2H, sp)
def act(): rt.cudaMemcpyAsync(da.data_ptr(), ha.data_ptr(), N * 22 * 4, H2D, sp)
cases = {"H2D actions 360 KB": act, "D2H full obs 2.46 MB contiguous": full, "D2H window part 4096 x 500 B (pitch 600)": window,
         "D2H self part 4096 x 100 B (pitch 600)": selfp, "D2H reward+done+info (3 copies, 151 KB)": small,
         "self part + 3 small": lambda: (selfp(), small()), "window + self + small (all pitched)": lambda: (window(), selfp(), small())}
for with_busy in (False, True):
    for name, fn in cases.items():
        ts = []
        for k in range(60):
            torch.cuda.synchronize()
            if with_busy: busy()
            t0 = time.perf_counter(); fn(); s.synchronize(); ts.append(time.perf_counter() - t0)
            torch.cuda.synchronize()
        ts = np.array(ts[10:]) * 1e6
        print(f"{'[GEMM running] ' if with_busy else ''}{name:50s} median {np.median(ts):7.1f} us  min {ts.min():7.1f} us", flush=True)
assert torch.equal(h, d.cpu())
