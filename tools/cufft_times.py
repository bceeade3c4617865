"""cuFFT (through torch.fft on CUDA) r2c / c2r times at the benchmark sizes, for the comparison table of DESIGN.md section 6."""
import torch

for shape, T in (((8192, 8192), torch.float64), ((4096, 4096), torch.float64), ((8192, 8192), torch.float32), ((4096, 4096), torch.float32),
                 ((512, 512, 512), torch.float64), ((512, 512, 512), torch.float32), ((1024, 1024, 1024), torch.float32), ((256, 2048, 2048), torch.float32)):
    x = torch.randn(shape, dtype=T, device="cuda")
    dims = tuple(range(len(shape)))
    xh = torch.fft.rfftn(x, dim=dims)
    res = {}
    for name, fn in (("r2c", lambda: torch.fft.rfftn(x, dim=dims)), ("c2r", lambda: torch.fft.irfftn(xh, s=shape, dim=dims))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            fn()
        b.record()
        torch.cuda.synchronize()
        res[name] = a.elapsed_time(b) / 10
    print(f"cufft {'x'.join(map(str, shape))} {str(T).split('.')[-1]}: r2c {res['r2c']:.3f} ms  c2r {res['c2r']:.3f} ms (torch.fft.irfftn includes its own normalisation pass)")
    del x, xh
    torch.cuda.empty_cache()
