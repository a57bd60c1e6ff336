import os, sys, time, threading
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from video_codec_pipeline_b200 import api
bench.select_workload("1080p", -1)
frames = bench.make_workload(32); n, fb = frames.shape
p = api.default_params(1920, 1080, fps=30, gop=60, qp_i=25, qp_p=27, slices=1)
host = torch.from_numpy(frames).pin_memory()
def run(nthreads, mode, steps=4):
    ses = [api.Session(p, n) for _ in range(nthreads)]
    outs = [torch.empty(n * fb // 2 + (1 << 20), dtype=torch.uint8).pin_memory().numpy() for _ in range(nthreads)]
    for s in ses: s.upload(host.data_ptr(), n); s.encode()
    def w(i, cnt):
        for _ in range(cnt):
            if mode in ("all", "up"): ses[i].upload(host.data_ptr(), n)
            if mode in ("all", "enc"): ses[i].encode()
            if mode == "all": ses[i].download(out=outs[i])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    th = [threading.Thread(target=w, args=(i, steps)) for i in range(nthreads)]
    [t.start() for t in th]; [t.join() for t in th]
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    for s in ses: s.close()
    print("threads", nthreads, mode, "ms/step %.1f" % (1000 * dt / (steps * nthreads)), "fps %.0f" % (n * steps * nthreads / dt), flush=True)
for nt, mode in ((1, "enc"), (2, "enc"), (3, "enc"), (1, "up"), (2, "up"), (1, "all"), (2, "all")):
    run(nt, mode)
