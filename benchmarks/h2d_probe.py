#!/usr/bin/env python
"""Host-to-device bandwidth of one pinned buffer split over k concurrent CUDA streams (what bounds bench.py's e2e)."""
import torch

def main():
    n = 26_600_000
    h = torch.randint(0, 255, (n,), dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for k in (1, 2, 3, 4, 6, 8, 12):
        streams = [torch.cuda.Stream() for _ in range(k)]
        step = (n + k - 1) // k
        def go():
            ev = torch.cuda.current_stream().record_event()
            for i, st in enumerate(streams):
                st.wait_event(ev)
                with torch.cuda.stream(st):
                    d[i * step:(i + 1) * step].copy_(h[i * step:(i + 1) * step], non_blocking=True)
                torch.cuda.current_stream().wait_event(st.record_event())
        for _ in range(3):
            go()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            go()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 20
        print("H2D 26.6 MB over %2d streams: %.4f ms  %.1f GB/s" % (k, ms, n / ms / 1e6))

if __name__ == "__main__":
    main()


def wc_probe():
    """Same copy from a write-combined pinned buffer (cudaHostAlloc flag 4) -- does the platform care?"""
    import ctypes
    import numpy as np
    rt = ctypes.CDLL("libcudart.so")
    n = 26_600_000
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), ctypes.c_uint(4))
    if rc != 0:
        print("cudaHostAlloc(WC) failed", rc)
        return
    arr = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(n,))
    arr[:] = 7
    h = torch.from_numpy(arr)
    print("write-combined buffer is_pinned:", h.is_pinned())
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for k in (1, 2):
        streams = [torch.cuda.Stream() for _ in range(k)]
        step = (n + k - 1) // k

        def go():
            ev = torch.cuda.current_stream().record_event()
            for i, st in enumerate(streams):
                st.wait_event(ev)
                with torch.cuda.stream(st):
                    d[i * step:(i + 1) * step].copy_(h[i * step:(i + 1) * step], non_blocking=True)
                torch.cuda.current_stream().wait_event(st.record_event())
        for _ in range(3):
            go()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            go()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 20
        print("WC H2D 26.6 MB over %2d streams: %.4f ms  %.1f GB/s" % (k, ms, n / ms / 1e6))
    assert int(d[12345]) == 7
    rt.cudaFreeHost(p)


if __name__ == "__main__" and "--wc" in __import__("sys").argv:
    wc_probe()
