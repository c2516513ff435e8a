#!/usr/bin/env python
"""Device-to-host copy bandwidth per GPU, alone and all GPUs at once, into (a) cudaHostAlloc'ed buffers and
(b) cudaHostRegister'ed pages of a /dev/shm mapping (what the multi-rank bench shares its result arrays through)."""
import json
import os
import threading
import time

import numpy as np
import torch

G = torch.cuda.device_count()
GB = 4
n = GB << 30
out = {"gpus": G, "bytes_per_copy": n}
src = [torch.empty(n, dtype=torch.uint8, device=f"cuda:{g}") for g in range(G)]
pinned = [torch.empty(n, dtype=torch.uint8, pin_memory=True) for g in range(G)]
path = f"/dev/shm/capsb_probe_{os.getpid()}"
shm = np.memmap(path, dtype=np.uint8, mode="w+", shape=(G * n,))
shm[:] = 0
os.unlink(path)
rc = torch.cuda.cudart().cudaHostRegister(shm.ctypes.data, shm.nbytes, 0)
shm_t = [torch.from_numpy(shm[g * n:(g + 1) * n]) for g in range(G)]


def run(dst, gpus):
    res = {}

    def one(g):
        torch.cuda.set_device(g)
        s = torch.cuda.Stream(device=g)
        with torch.cuda.stream(s):
            dst[g].copy_(src[g], non_blocking=True)  # warm-up
            s.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                dst[g].copy_(src[g], non_blocking=True)
            s.synchronize()
            res[g] = 3 * n / (time.perf_counter() - t0) / 1e9

    ts = [threading.Thread(target=one, args=(g,)) for g in gpus]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return {str(g): round(v, 1) for g, v in res.items()}


out["cudaHostAlloc_alone_GBps"] = run(pinned, [0])
out["cudaHostAlloc_all_GBps"] = run(pinned, list(range(G)))
out["shm_registered_rc"] = int(rc)
out["shm_registered_alone_GBps"] = run(shm_t, [0])
out["shm_registered_all_GBps"] = run(shm_t, list(range(G)))
print(json.dumps(out), flush=True)
