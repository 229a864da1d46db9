"""Pinned host -> device copy rate of one bench batch (403 MB) for the three kinds of pinned memory a caller could hand to
mmw_submit_host: cudaHostAlloc default, write-combined, and torch's pin_memory.  Run on the GPU box."""
import ctypes as C
import time

import torch

rt = C.CDLL("libcudart.so")
N = 402653184
dev = torch.empty(N, dtype=torch.uint8, device="cuda")
stream = torch.cuda.Stream()


def rate(host_ptr, label):
    for _ in range(3):
        rt.cudaMemcpyAsync(C.c_void_p(dev.data_ptr()), C.c_void_p(host_ptr), C.c_size_t(N), 1, C.c_void_p(stream.cuda_stream))
    rt.cudaStreamSynchronize(C.c_void_p(stream.cuda_stream))
    t0 = time.perf_counter()
    K = 20
    for _ in range(K):
        rt.cudaMemcpyAsync(C.c_void_p(dev.data_ptr()), C.c_void_p(host_ptr), C.c_size_t(N), 1, C.c_void_p(stream.cuda_stream))
    rt.cudaStreamSynchronize(C.c_void_p(stream.cuda_stream))
    dt = time.perf_counter() - t0
    print(f"{label}: {N * K / dt / 1e9:.2f} GB/s", flush=True)


for flags, label in ((0, "cudaHostAlloc default"), (4, "cudaHostAlloc write-combined"), (1, "cudaHostAlloc portable")):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(N), flags) == 0
    C.memset(p, 1, N)
    rate(p.value, label)
    rt.cudaFreeHost(p)
t = torch.empty(N, dtype=torch.uint8, pin_memory=True)
t.fill_(1)
rate(t.data_ptr(), "torch pin_memory")
