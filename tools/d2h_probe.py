"""Device -> pinned-host copy bandwidth on this box: one large copy against 2 / 4 concurrent streams and several chunk sizes,
and the host's own memcpy rate on T threads (what the host expansion of cmg_set_host_expand is bound by).
    python tools/d2h_probe.py [GiB]"""
import ctypes, os, sys, threading, time
import numpy as np
import torch

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
n = int(gib * (1 << 30)) // 8
dev = torch.empty(n, dtype=torch.float64, device="cuda").fill_(1.0)
t0 = time.time()
host = torch.empty(n, dtype=torch.float64, pin_memory=True)
print("pinned alloc of %.1f GiB: %.2f s" % (gib, time.time() - t0), flush=True)
host.fill_(0.0)

def run(nstreams, chunk_mib):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    chunk = chunk_mib * (1 << 20) // 8
    torch.cuda.synchronize()
    t = time.perf_counter()
    k = 0
    for off in range(0, n, chunk):
        with torch.cuda.stream(streams[k % nstreams]):
            host[off:off + chunk].copy_(dev[off:off + chunk], non_blocking=True)
        k += 1
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print("D2H %d stream(s), chunk %5d MiB: %.1f GB/s" % (nstreams, chunk_mib, n * 8 / dt / 1e9), flush=True)

for ns, ch in ((1, 1 << 14), (1, 1024), (1, 64), (2, 1024), (2, 64), (4, 256), (4, 16)):
    run(ns, min(ch, int(gib * 1024)))

# host memcpy on T threads (numpy releases the GIL in copyto for large contiguous blocks)
src = host.numpy()
dst = np.empty_like(src)
dst.fill(0)
for T in sorted({1, 2, 4, 8, 16, 32, os.cpu_count() or 1}):
    if T > (os.cpu_count() or 1):
        continue
    parts = np.array_split(np.arange(0, n + 1, max(1, n // (T * 8)))[:], 1)[0]
    bounds = np.linspace(0, n, T + 1).astype(np.int64)
    def work(a, b):
        np.copyto(dst[a:b], src[a:b])
    th = [threading.Thread(target=work, args=(bounds[i], bounds[i + 1])) for i in range(T)]
    t = time.perf_counter()
    [x.start() for x in th]; [x.join() for x in th]
    dt = time.perf_counter() - t
    print("host memcpy %2d threads: %.1f GB/s copied (read + write = 2x)" % (T, n * 8 / dt / 1e9), flush=True)
