"""Where the time of the plugin call goes: vcpenc_transcode(y4m on /dev/shm -> mp4) with VCPENC_TRACE=1, one and two
worker threads.  python tools/transcode_probe.py [gops] (needs the GPU; knobs through the environment)"""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_codec_pipeline_b200 import api, synth  # noqa: E402

gops = int(sys.argv[1]) if len(sys.argv) > 1 else 32
w, h, fps, gop = 1920, 1080, 30, 60
two = np.concatenate([synth.make_clip(w, h, gop, seed=1080), synth.make_clip(w, h, gop, seed=1081, start=gop)])
src = "/dev/shm/vcp_probe_%d.y4m" % os.getpid()
with open(src, "wb") as f:
    f.write(b"YUV4MPEG2 W%d H%d F%d:1 Ip A1:1 C420\n" % (w, h, fps))
    for g in range(gops):
        for i in range(gop):
            f.write(b"FRAME\n")
            f.write(memoryview(two[(g % 2) * gop + i]))
tokens = "-c:v libx264 -profile:v baseline -coder 0 -g 60 -qp 27 -slices 1"


def task(i, reps):
    api.set_thread_device(0)
    for _ in range(reps):
        api.transcode(src, "/dev/shm/vcp_probe_%d_%d.mp4" % (os.getpid(), i), tokens)


for nthr in (1, 2, 3):
    th = [threading.Thread(target=task, args=(i, 1)) for i in range(nthr)]      # warm-up: sessions, pinned buffers
    [t.start() for t in th]; [t.join() for t in th]
    os.environ["VCPENC_TRACE"] = "1"
    t0 = time.perf_counter()
    th = [threading.Thread(target=task, args=(i, 2)) for i in range(nthr)]
    [t.start() for t in th]; [t.join() for t in th]
    dt = time.perf_counter() - t0
    os.environ.pop("VCPENC_TRACE", None)
    print("threads %d: %.0f fps (%d tasks of %d frames in %.3f s)" % (nthr, nthr * 2 * gops * gop / dt, nthr * 2, gops * gop, dt), flush=True)
for i in range(3):
    try:
        os.remove("/dev/shm/vcp_probe_%d_%d.mp4" % (os.getpid(), i))
    except OSError:
        pass
os.remove(src)
