"""Probe: weight-gradient GEMM forms under cuBLAS 12.9 FP32 emulation."""
import ctypes, os, sys
os.environ["CUBLAS_EMULATE_SINGLE_PRECISION"] = "1"
for lib in ("libcublasLt.so.12", "libcublas.so.12"):
    ctypes.CDLL(os.path.join("/usr/local/cuda/lib64", lib), mode=ctypes.RTLD_GLOBAL)
import torch
torch.backends.cuda.matmul.allow_tf32 = False
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
B = 65536
for kin, nout in [(512, 256), (256, 128), (13, 512), (256, 1), (480, 1024), (1024, 1024), (1024, 512), (128, 512)]:
    x = torch.randn(B, kin, device="cuda"); g = torch.randn(B, nout, device="cuda")
    a = t(lambda: x.t() @ g)
    b = t(lambda: (g.t() @ x).t())
    xc = x.t().contiguous()
    c = t(lambda: xc @ g)
    S = 16
    d = t(lambda: torch.bmm(x.view(S, B // S, kin).transpose(1, 2), g.view(S, B // S, nout)).sum(0))
    print(f"dW {kin}x{nout}: x.t()@g {a:.3f} ms | (g.t()@x).t() {b:.3f} | contiguous xT @ g {c:.3f} | bmm split-K16 {d:.3f}   ideal@150TF {2*B*kin*nout/150e12*1e3:.3f}")
for kin, nout in [(512, 256), (256, 128), (480, 1024)]:
    g = torch.randn(B, nout, device="cuda"); w = torch.randn(kin, nout, device="cuda")
    a = t(lambda: g @ w.t())
    print(f"dX {nout}->{kin}: g@w.t() {a:.3f} ms  ideal {2*B*kin*nout/150e12*1e3:.3f}")
