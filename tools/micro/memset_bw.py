import torch, time
x = torch.empty(8 * 1024**3, dtype=torch.uint8, device="cuda")
y = torch.empty(4 * 1024**3, dtype=torch.uint8, device="cuda")
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = t(lambda: x.zero_())
print("memset 8 GiB (torch zero_):", ms, "ms", x.numel() / ms / 1e6, "GB/s")
ms = t(lambda: x.fill_(7))
print("fill 8 GiB:", ms, "ms", x.numel() / ms / 1e6, "GB/s")
xd = x.view(torch.float64)
ms = t(lambda: xd.fill_(1.5))
print("fill f64 8 GiB:", ms, "ms", x.numel() / ms / 1e6, "GB/s")
ms = t(lambda: y.copy_(x[:y.numel()]))
print("copy 4 GiB (r+w 8 GiB):", ms, "ms", 2 * y.numel() / ms / 1e6, "GB/s")
import ctypes
cudart = ctypes.CDLL("libcudart.so")
ms = t(lambda: cudart.cudaMemsetAsync(ctypes.c_void_p(x.data_ptr()), 0, ctypes.c_size_t(x.numel()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
print("cudaMemsetAsync 8 GiB:", ms, "ms", x.numel() / ms / 1e6, "GB/s")
