"""Probe: torch fp32 GEMM through the image's cuBLAS 12.9 with FP32 emulation (BF16x9) vs native."""
import ctypes, os, sys, time
mode = sys.argv[1] if len(sys.argv) > 1 else "native"
if mode != "native":
    os.environ["CUBLAS_EMULATE_SINGLE_PRECISION"] = "1"
    if mode == "eager":
        os.environ["CUBLAS_EMULATION_STRATEGY"] = "eager"
    for lib in ("libcublasLt.so.12", "libcublas.so.12"):
        ctypes.CDLL(os.path.join("/usr/local/cuda/lib64", lib), mode=ctypes.RTLD_GLOBAL)
import torch
torch.backends.cuda.matmul.allow_tf32 = False
maps = [l.split()[-1] for l in open("/proc/self/maps") if "cublas" in l]
print(mode, "cublas loaded from:", sorted(set(maps)))
def bench(M, N, K, bias=False):
    a = torch.randn(M, K, device="cuda"); b = torch.randn(K, N, device="cuda")
    for _ in range(3): c = a @ b
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): c = a @ b
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    ref = (a[:256].double() @ b.double())
    err = ((c[:256].double() - ref).abs().max() / ref.abs().max()).item()
    print(f"  {M}x{N}x{K}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s  max err/max|ref| = {err:.2e}")
for shp in [(65536, 1024, 1024), (65536, 1024, 479), (65536, 512, 1024), (65536, 256, 512), (65536, 512, 13), (8192, 8192, 8192)]:
    bench(*shp)
lin = torch.nn.Linear(1024, 1024).cuda(); x = torch.randn(65536, 1024, device="cuda", requires_grad=True)
for _ in range(3): lin(x).sum().backward()
torch.cuda.synchronize(); t=time.time()
for _ in range(10): lin(x).sum().backward()
torch.cuda.synchronize(); print(f"  Linear(1024,1024) fwd+bwd B=65536: {(time.time()-t)*100:.2f} ms")
