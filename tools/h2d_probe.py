"""Pinned-host -> device copy bandwidth with 1, 2, 4 concurrent copy streams (does a second DMA engine add anything on this link?)."""
import time, torch
GB = 4
h = torch.empty(GB << 30, dtype=torch.uint8, pin_memory=True); h.fill_(1)
d = torch.empty(GB << 30, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
for ns in (1, 2, 4, 1, 2):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    part = (GB << 30) // ns
    best = 0
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                d[i * part:(i + 1) * part].copy_(h[i * part:(i + 1) * part], non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        best = max(best, (GB << 30) / dt / 1e9)
    print("%d stream(s): %.2f GB/s" % (ns, best))
# chunked: 256 MiB pieces back to back on one stream vs alternating two streams
for ns in (1, 2):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    piece = 256 << 20
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range((GB << 30) // piece):
        with torch.cuda.stream(streams[k % ns]):
            d[k * piece:(k + 1) * piece].copy_(h[k * piece:(k + 1) * piece], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("256 MiB pieces on %d stream(s): %.2f GB/s" % (ns, (GB << 30) / dt / 1e9))
