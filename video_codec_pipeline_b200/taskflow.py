"""Task flow around the executor: a re-statement of the VCP consumer's `processTask` and of the
Redis-Streams calls it makes, so that BASELINE.json configs #1 / #5 (task -> Redis stream ->
consumer -> encoded file) can run in an image that has neither Go nor Redis.

STAND-INS, clearly: the reference's control plane is Go (`cmd/consumer.go`, `internal/redis/
stream.go`) and stays the reference's; nothing here is meant to replace it.  What this module
keeps faithful is the CONTRACT the executor lives in:

  * wire format of a task            internal/redis/stream.go:127-137 (XADD fields), :180-216 (parse)
  * stream / group names             internal/redis/stream.go:13-14  ("vcp:tasks", "gpu_encoders")
  * ReadGroup / Acknowledge          internal/redis/stream.go:143-164, 219-227 (XREADGROUP ">" COUNT 1 BLOCK 3000;
                                     XACK then XDEL)
  * validity gate                    cmd/consumer.go:136-142 (empty ID / InputPath -> ACK and drop)
  * processTask                      cmd/consumer.go:220-318: wait for a stable input (4 stats 500 ms apart,
                                     :321-367) -> mkdir -> encode (60 min) -> optional verify -> delete the
                                     input BEFORE the ACK -> ACK; every failure: remove output, ACK, drop
  * worker pool                      cmd/consumer.go:119-175: `-j` workers, channel of depth 2j fed by a reader
  * builtin presets                  internal/config/config.go:44-52

`MiniRedis` is a tiny in-process RESP2 server that speaks exactly the commands above (PING,
XGROUP CREATE ... MKSTREAM, XADD, XREADGROUP, XACK, XDEL, XLEN, XPENDING) — enough for go-redis'
calls in the reference and for the client below; a real Redis works the same way.
"""
from __future__ import annotations

import os
import sys
import queue
import socket
import socketserver
import threading
import time
from dataclasses import dataclass, field

STREAM = "vcp:tasks"            # internal/redis/stream.go:13
GROUP = "gpu_encoders"          # internal/redis/stream.go:14

# internal/config/config.go:44-52
PRESETS = {
    "h264-nvenc": "-c:v h264_nvenc -preset p4 -b:v 10M -c:a aac -b:a 128k -movflags +faststart",
    "h264-nvenc-hq": "-c:v h264_nvenc -preset p7 -tune hq -b:v 15M -maxrate 20M -bufsize 30M -c:a aac -b:a 192k -movflags +faststart",
    "h265-nvenc": "-c:v hevc_nvenc -preset p4 -b:v 8M -c:a aac -b:a 128k -movflags +faststart",
    "h265-nvenc-hq": "-c:v hevc_nvenc -preset p7 -tune hq -b:v 10M -c:a aac -b:a 192k -movflags +faststart",
    "h264-cpu": "-c:v libx264 -preset medium -crf 23 -c:a aac -b:a 128k -movflags +faststart",
    "h265-cpu": "-c:v libx265 -preset medium -crf 28 -c:a aac -b:a 128k -movflags +faststart",
    "copy": "-c copy",
}


@dataclass
class Task:
    """internal/redis/stream.go:30-48."""
    id: str = ""
    input_path: str = ""
    original_name: str = ""
    output_dir: str = ""
    output_name: str = ""
    ffmpeg_args: str = ""
    verify_output: bool = False
    source_ip: str = ""
    retry: int = 0
    message_id: str = ""

    def fields(self):
        return ["task_id", self.id, "input_path", self.input_path, "original_name", self.original_name,
                "output_dir", self.output_dir, "output_name", self.output_name, "ffmpeg_args", self.ffmpeg_args,
                "verify_output", "true" if self.verify_output else "false", "source_ip", self.source_ip,
                "retry", str(self.retry)]

    @staticmethod
    def parse(message_id, kv):
        d = {kv[i]: kv[i + 1] for i in range(0, len(kv) - 1, 2)}
        t = Task(message_id=message_id)
        t.id = d.get("task_id", ""); t.input_path = d.get("input_path", "")
        t.original_name = d.get("original_name", ""); t.output_dir = d.get("output_dir", "")
        t.output_name = d.get("output_name", ""); t.ffmpeg_args = d.get("ffmpeg_args", "")
        t.verify_output = d.get("verify_output", "") in ("true", "1")       # stream.go:204
        t.source_ip = d.get("source_ip", "")
        try:
            t.retry = int(d.get("retry", "0"))
        except ValueError:
            t.retry = 0
        return t


# ---- RESP2 ----------------------------------------------------------------------------------------
def _enc(x) -> bytes:
    if x is None:
        return b"$-1\r\n"
    if isinstance(x, int):
        return b":%d\r\n" % x
    if isinstance(x, (list, tuple)):
        return b"*%d\r\n" % len(x) + b"".join(_enc(e) for e in x)
    if isinstance(x, Exception):
        return b"-" + str(x).encode() + b"\r\n"
    if isinstance(x, str):
        x = x.encode()
    return b"$%d\r\n%s\r\n" % (len(x), x)


def _read(f):
    line = f.readline()
    if not line:
        raise EOFError
    t, rest = line[:1], line[1:-2]
    if t == b"+":
        return rest.decode()
    if t == b"-":
        raise RuntimeError(rest.decode())
    if t == b":":
        return int(rest)
    if t == b"$":
        n = int(rest)
        if n < 0:
            return None
        data = f.read(n + 2)[:-2]
        return data.decode()
    if t == b"*":
        n = int(rest)
        return None if n < 0 else [_read(f) for _ in range(n)]
    raise RuntimeError("bad RESP type %r" % t)


class _Stream:
    def __init__(self):
        self.entries = []          # [(id, [k, v, ...])]
        self.groups = {}           # group -> {"next": index into entries, "pending": {id: consumer}}
        self.seq = 0


class MiniRedis:
    """In-process stand-in for the Redis server of configs #1/#5 (Streams subset, RESP2)."""

    def __init__(self, host="127.0.0.1", port=0):
        self.streams = {}
        self.cv = threading.Condition()
        outer = self

        class H(socketserver.StreamRequestHandler):
            def handle(self):
                while True:
                    try:
                        cmd = _read(self.rfile)
                    except (EOFError, ConnectionError):
                        return
                    try:
                        rep = outer.execute(cmd)
                        self.wfile.write(b"+OK\r\n" if rep == "OK" else (b"+PONG\r\n" if rep == "PONG" else _enc(rep)))
                    except Exception as e:  # noqa: BLE001
                        self.wfile.write(_enc(e))

        class S(socketserver.ThreadingTCPServer):
            allow_reuse_address = True
            daemon_threads = True

        self.server = S((host, port), H)
        self.addr = self.server.server_address
        self.thread = threading.Thread(target=self.server.serve_forever, daemon=True)
        self.thread.start()

    def close(self):
        self.server.shutdown()
        self.server.server_close()

    def execute(self, cmd):
        op = cmd[0].upper()
        a = cmd[1:]
        with self.cv:
            if op == "PING":
                return "PONG"
            if op == "XGROUP" and a[0].upper() == "CREATE":
                st = self.streams.get(a[1])
                if st is None:
                    if not any(x.upper() == "MKSTREAM" for x in a[4:]):
                        raise RuntimeError("ERR no such key")
                    st = self.streams[a[1]] = _Stream()
                if a[2] in st.groups:
                    raise RuntimeError("BUSYGROUP Consumer Group name already exists")
                st.groups[a[2]] = {"next": 0 if a[3] == "0" else len(st.entries), "pending": {}}
                return "OK"
            if op == "XADD":
                st = self.streams.setdefault(a[0], _Stream())
                st.seq += 1
                mid = "%d-%d" % (int(time.time() * 1000), st.seq)
                st.entries.append((mid, list(a[2:])))
                self.cv.notify_all()
                return mid
            if op == "XLEN":
                st = self.streams.get(a[0])
                return len(st.entries) if st else 0
            if op == "XREADGROUP":
                group, consumer = a[1], a[2]
                count, block, i = 1 << 30, None, 3
                while a[i].upper() != "STREAMS":
                    if a[i].upper() == "COUNT":
                        count = int(a[i + 1])
                    elif a[i].upper() == "BLOCK":
                        block = int(a[i + 1])
                    i += 2
                key, start = a[i + 1], a[i + 2]
                deadline = None if block is None else time.time() + block / 1000.0
                while True:
                    st = self.streams.get(key)
                    if st is None or group not in st.groups:
                        raise RuntimeError("NOGROUP No such key or consumer group")
                    g = st.groups[group]
                    if start == ">":
                        got = []
                        while g["next"] < len(st.entries) and len(got) < count:
                            e = st.entries[g["next"]]
                            g["next"] += 1
                            if e is not None:
                                g["pending"][e[0]] = consumer
                                got.append(e)
                    else:       # history of this consumer's pending entries
                        got = [e for e in st.entries if e is not None and g["pending"].get(e[0]) == consumer][:count]
                    if got or start != ">" or block is None:
                        return [[key, [[e[0], e[1]] for e in got]]] if got else None
                    left = deadline - time.time()
                    if left <= 0:
                        return None
                    self.cv.wait(left)
            if op == "XACK":
                st = self.streams.get(a[0])
                n = 0
                if st and a[1] in st.groups:
                    for mid in a[2:]:
                        n += st.groups[a[1]]["pending"].pop(mid, None) is not None
                return n
            if op == "XDEL":
                st = self.streams.get(a[0])
                n = 0
                if st:
                    for k, e in enumerate(st.entries):
                        if e is not None and e[0] in a[1:]:
                            st.entries[k] = None
                            n += 1
                return n
            if op == "XPENDING":
                st = self.streams.get(a[0])
                p = st.groups[a[1]]["pending"] if st and a[1] in st.groups else {}
                if not p:
                    return [0, None, None, None]
                ids = sorted(p)
                per = {}
                for c in p.values():
                    per[c] = per.get(c, 0) + 1
                return [len(p), ids[0], ids[-1], [[c, str(n)] for c, n in per.items()]]
            raise RuntimeError("ERR unknown command '%s'" % op)


class RedisClient:
    """Minimal RESP2 client: one connection, one command at a time (what one consumer needs)."""

    def __init__(self, host="127.0.0.1", port=6379):
        self.sock = socket.create_connection((host, port))
        self.f = self.sock.makefile("rb")
        self.lock = threading.Lock()

    def call(self, *args):
        with self.lock:
            self.sock.sendall(_enc([str(a) for a in args]))
            return _read(self.f)

    def close(self):
        try:
            self.sock.close()
        except OSError:
            pass

    # the calls of internal/redis/stream.go
    def ping(self):
        return self.call("PING")

    def create_consumer_group(self, stream=STREAM, group=GROUP):
        try:
            self.call("XGROUP", "CREATE", stream, group, "0", "MKSTREAM")
        except RuntimeError as e:
            if "BUSYGROUP" not in str(e):
                raise

    def publish(self, task: Task) -> str:
        return self.call("XADD", STREAM, "*", *task.fields())

    def read_group(self, consumer, count=1, block_ms=3000, group=GROUP):
        rep = self.call("XREADGROUP", "GROUP", group, consumer, "COUNT", count, "BLOCK", block_ms, "STREAMS", STREAM, ">")
        out = []
        for _key, msgs in rep or []:
            for mid, kv in msgs:
                out.append(Task.parse(mid, kv))
        return out

    def acknowledge(self, message_id, group=GROUP):
        self.call("XACK", STREAM, group, message_id)     # ACK first,
        self.call("XDEL", STREAM, message_id)            # then free the entry (stream.go:219-227)


# ---- processTask -------------------------------------------------------------------------------
def wait_for_file(path, timeout=30.0, poll=0.5, cancelled=lambda: False):
    """cmd/consumer.go:321-367: the size must be > 0 and unchanged on 3 consecutive polls after the
    first sighting (4 stats), then one open() proves it is readable."""
    if not path:
        raise RuntimeError("文件路径为空")
    deadline = time.time() + timeout
    last, stable = -1, 0
    while time.time() < deadline:
        if cancelled():
            raise RuntimeError("context canceled")
        try:
            size = os.stat(path).st_size
        except FileNotFoundError:
            time.sleep(poll)
            continue
        if size > 0:
            if size == last:
                stable += 1
                if stable >= 3:
                    with open(path, "rb"):
                        return
            else:
                stable, last = 0, size
        time.sleep(poll)
    raise RuntimeError("等待文件超时")


@dataclass
class Stats:
    processed: int = 0
    success: int = 0
    failed: int = 0
    log: list = field(default_factory=list)


def process_task(client, task: Task, transcode, verify, poll=0.5, cancelled=lambda: False, stats: Stats | None = None):
    """cmd/consumer.go:220-318.  `transcode(input, output, ffmpeg_args, timeout_ms)` and `verify(path)`
    are the two executor calls (api.transcode / api.verify for the B200 path); they raise on failure."""
    t_start = time.time()
    marks = {}

    def note(kind, msg=""):
        if stats is not None:
            stats.log.append((task.id, kind, msg, dict(marks, total=round(time.time() - t_start, 3))))

    def fail(reason, remove=None):
        note("failed", reason)
        if remove:
            try:
                os.remove(remove)
            except OSError:
                pass
        if task.message_id:
            client.acknowledge(task.message_id)          # failed tasks are ACKed and dropped
        return False

    if cancelled():
        return fail("context_cancelled")
    try:
        wait_for_file(task.input_path, 30.0, poll, cancelled)
    except Exception as e:  # noqa: BLE001
        return fail("input_file_unavailable: %s" % e)
    marks["wait"] = round(time.time() - t_start, 3)
    try:
        os.makedirs(task.output_dir, mode=0o755, exist_ok=True)
    except OSError as e:
        return fail("mkdir_failed: %s" % e)
    out = os.path.join(task.output_dir, task.output_name)
    try:
        transcode(task.input_path, out, task.ffmpeg_args, 60 * 60 * 1000)
    except Exception as e:  # noqa: BLE001
        return fail("ffmpeg_failed: %s" % e, remove=out)
    marks["encoded"] = round(time.time() - t_start, 3)
    if task.verify_output:
        try:
            verify(out)
        except Exception as e:  # noqa: BLE001
            return fail("verify_failed: %s" % e, remove=out)
    try:
        os.remove(task.input_path)                        # the source goes BEFORE the ACK (:288 vs :301)
    except OSError as e:
        note("warn", "delete_input_file_failed %s" % e)
    if task.message_id:
        try:
            client.acknowledge(task.message_id)
        except Exception as e:  # noqa: BLE001
            note("failed", "task_ack_failed %s" % e)
            return False
    note("success", out)
    return True


class Consumer:
    """cmd/consumer.go:119-175: one reader (COUNT 1, BLOCK 3 s) feeding a channel of depth 2j, j workers."""

    def __init__(self, addr, name, transcode, verify, concurrency=1, poll=0.5, on_exit=None):
        self.addr, self.name, self.j, self.poll = addr, name, max(1, concurrency), poll
        self.on_exit = on_exit
        self.transcode, self.verify = transcode, verify
        self.stats = Stats()
        self.stop = threading.Event()
        self.ch = queue.Queue(maxsize=2 * self.j)
        self.threads = []
        self.lock = threading.Lock()

    def _worker(self):
        client = RedisClient(*self.addr)
        while not self.stop.is_set():
            try:
                task = self.ch.get(timeout=0.1)
            except queue.Empty:
                continue
            if not task.id or not task.input_path:            # validity gate (:136-142)
                if task.message_id:
                    client.acknowledge(task.message_id)
                continue
            ok = process_task(client, task, self.transcode, self.verify, self.poll, self.stop.is_set, self.stats)
            with self.lock:
                self.stats.processed += 1
                self.stats.success += ok
                self.stats.failed += not ok
        client.close()
        if self.on_exit:
            self.on_exit()          # e.g. api.thread_release: free the worker's cached encoder session

    def _reader(self):
        client = RedisClient(*self.addr)
        client.create_consumer_group()
        while not self.stop.is_set():
            try:
                tasks = client.read_group(self.name, 1, 200)
            except Exception:  # noqa: BLE001
                time.sleep(0.2)
                continue
            for t in tasks:
                while not self.stop.is_set():
                    try:
                        self.ch.put(t, timeout=0.1)
                        break
                    except queue.Full:
                        pass
        client.close()

    def start(self):
        self.threads = [threading.Thread(target=self._worker, daemon=True) for _ in range(self.j)]
        self.threads.append(threading.Thread(target=self._reader, daemon=True))
        for t in self.threads:
            t.start()
        return self

    def shutdown(self, wait=5.0):
        self.stop.set()
        for t in self.threads:
            t.join(wait)


def main(argv=None):
    """`python -m video_codec_pipeline_b200.taskflow --clips N --consumers K -j J`: config #5 in miniature —
    N synthetic clips (mixed 720p/1080p) through MiniRedis into K consumers (one per visible GPU)."""
    import argparse
    import json
    import tempfile

    import numpy as np

    from . import api, synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=8)
    ap.add_argument("--frames", type=int, default=60)
    ap.add_argument("--consumers", type=int, default=0, help="0 = one per visible GPU")
    ap.add_argument("-j", type=int, default=1)
    ap.add_argument("--preset", default="h264-cpu")
    ap.add_argument("--poll", type=float, default=0.5, help="input stability poll, s (reference: 0.5)")
    a = ap.parse_args(argv)
    ngpu = api.device_count()
    if ngpu < 1:
        raise SystemExit("no CUDA device: the executor has no CPU fallback")
    k = a.consumers or ngpu
    srv = MiniRedis()
    tmp = tempfile.mkdtemp(prefix="vcpflow_")
    sizes = [(1280, 720), (1920, 1080)]
    prod = RedisClient(*srv.addr)
    prod.create_consumer_group()
    total_frames = 0
    for i in range(a.clips):
        w, h = sizes[i % len(sizes)]
        clip = synth.make_clip(w, h, a.frames, seed=5000 + i)
        path = os.path.join(tmp, "clip%03d.y4m" % i)
        with open(path, "wb") as f:
            f.write(b"YUV4MPEG2 W%d H%d F30:1 Ip A1:1 C420jpeg\n" % (w, h))
            for fr in clip:
                f.write(b"FRAME\n" + np.ascontiguousarray(fr).tobytes())
        total_frames += a.frames
        prod.publish(Task(id="task-%03d" % i, input_path=path, original_name=os.path.basename(path),
                          output_dir=os.path.join(tmp, "out"), output_name="clip%03d.mp4" % i,
                          ffmpeg_args=PRESETS[a.preset], verify_output=True))

    def make_exec(dev):
        def run(inp, out, args, timeout_ms):
            api.set_thread_device(dev)
            api.transcode(inp, out, args, timeout_ms)
        return run

    t0 = time.time()
    cons = [Consumer(srv.addr, "gpu%d" % (i % ngpu), make_exec(i % ngpu), api.verify, a.j, a.poll, api.thread_release).start() for i in range(k)]
    while sum(c.stats.processed for c in cons) < a.clips and time.time() - t0 < 3600:
        time.sleep(0.05)
    dt = time.time() - t0
    for c in cons:
        c.shutdown()
    ok = sum(c.stats.success for c in cons)
    if os.environ.get("VCPENC_TRACE"):
        for c in cons:
            for entry in c.stats.log:
                print("[taskflow]", c.name, entry, file=sys.stderr)
    print(json.dumps({"config": "configs[4]-style: %d clips x %d frames, preset %s, %d consumers x -j %d, MiniRedis stand-in" %
                      (a.clips, a.frames, a.preset, k, a.j), "tasks": a.clips, "succeeded": ok, "wall_s": round(dt, 3),
                      "tasks_per_s": round(a.clips / dt, 3), "aggregate_fps": round(total_frames / dt, 1),
                      "fixed_wait_per_task_s": round(3 * a.poll, 2)}))
    srv.close()
    return 0 if ok == a.clips else 1


if __name__ == "__main__":
    raise SystemExit(main())
