"""Developer tool (GPU box): host->device copy bandwidth from pinned memory, for the sizes bench.py's e2e leg
moves (one C3 logits tensor = 12.3 MB) -- alone, split over two streams, and for a large buffer."""
import torch

def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def main():
    dev = torch.device("cuda", 0)
    for mb in (12.3, 64, 256):
        n = int(mb * 1e6 / 4)
        h = torch.empty(n, dtype=torch.float32).pin_memory()
        d = torch.empty(n, dtype=torch.float32, device=dev)
        ms = timed(lambda: d.copy_(h, non_blocking=True))
        print("H2D %6.1f MB pinned, one stream : %.3f ms  %.1f GB/s" % (mb, ms, mb / ms))
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        half = n // 2
        def two():
            with torch.cuda.stream(s1):
                d[:half].copy_(h[:half], non_blocking=True)
            with torch.cuda.stream(s2):
                d[half:].copy_(h[half:], non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            two()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            two()
        s1.synchronize(); s2.synchronize()
        e1.record()
        torch.cuda.synchronize()
        # wall-clock style: events on the default stream do not bracket the side streams; use host timing instead
        import time
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            two()
        torch.cuda.synchronize()
        ms2 = (time.perf_counter() - t0) * 1e3 / 20
        print("H2D %6.1f MB pinned, two streams: %.3f ms  %.1f GB/s" % (mb, ms2, mb / ms2))
        hp = torch.empty(n, dtype=torch.float32)
        ms3 = timed(lambda: d.copy_(hp, non_blocking=True), n=5)
        print("H2D %6.1f MB pageable           : %.3f ms  %.1f GB/s" % (mb, ms3, mb / ms3))
    import subprocess
    print(subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current", "--format=csv"],
                         capture_output=True, text=True).stdout)

if __name__ == "__main__":
    main()
