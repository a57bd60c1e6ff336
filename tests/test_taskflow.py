"""Task flow around the executor: the re-stated processTask contract and the Redis
Streams subset it uses, with a fake executor on CPU and the real one on the GPU box."""
import os
import threading
import time

import numpy as np
import pytest

from video_codec_pipeline_b200 import taskflow as tf


@pytest.fixture()
def redis():
    srv = tf.MiniRedis()
    yield srv
    srv.close()


def test_resp_streams_subset(redis):
    c = tf.RedisClient(*redis.addr)
    assert c.ping() == "PONG"
    c.create_consumer_group()
    c.create_consumer_group()                                   # BUSYGROUP is swallowed (stream.go:109-113)
    t = tf.Task(id="a", input_path="/x/in.mp4", original_name="in.mp4", output_dir="/x/out", output_name="o.mp4",
                ffmpeg_args=tf.PRESETS["h264-cpu"], verify_output=True, source_ip="10.0.0.1", retry=2)
    mid = c.publish(t)
    assert c.call("XLEN", tf.STREAM) == 1
    got = c.read_group("gpu0", 1, 50)
    assert len(got) == 1 and got[0].message_id == mid
    g = got[0]
    assert (g.id, g.input_path, g.output_name, g.ffmpeg_args, g.verify_output, g.retry) == \
           ("a", "/x/in.mp4", "o.mp4", tf.PRESETS["h264-cpu"], True, 2)
    assert c.read_group("gpu1", 1, 20) == []                     # delivered to exactly one consumer
    assert c.call("XPENDING", tf.STREAM, tf.GROUP)[0] == 1
    c.acknowledge(mid)
    assert c.call("XPENDING", tf.STREAM, tf.GROUP)[0] == 0
    # blocking read wakes up on XADD
    res = []
    th = threading.Thread(target=lambda: res.extend(tf.RedisClient(*redis.addr).read_group("gpu0", 1, 2000)))
    th.start()
    time.sleep(0.1)
    c.publish(tf.Task(id="b", input_path="/y"))
    th.join(3)
    assert [r.id for r in res] == ["b"]
    assert tf.Task.parse("1-1", ["verify_output", "1"]).verify_output and not tf.Task.parse("1-1", ["verify_output", "yes"]).verify_output


def test_process_task_semantics(redis, tmp_path):
    c = tf.RedisClient(*redis.addr)
    c.create_consumer_group()
    calls = []

    def ok_exec(inp, out, args, timeout_ms):
        calls.append((inp, out, args, timeout_ms))
        open(out, "wb").write(b"encoded")

    def bad_exec(inp, out, args, timeout_ms):
        open(out, "wb").write(b"partial")
        raise RuntimeError("exit status 1")

    def mk(name):
        p = tmp_path / name
        p.write_bytes(b"x" * 100)
        t = tf.Task(id=name, input_path=str(p), original_name=name, output_dir=str(tmp_path / "out"), output_name=name + ".mp4",
                    ffmpeg_args="-c:v libx264 -crf 23", verify_output=True)
        c.publish(t)
        return c.read_group("w", 1, 50)[0]

    # success: output kept, input deleted before the ACK, message gone
    t = mk("a")
    assert tf.process_task(c, t, ok_exec, lambda p: None, poll=0.01)
    assert (tmp_path / "out" / "a.mp4").read_bytes() == b"encoded" and not os.path.exists(t.input_path)
    assert calls[0][2] == "-c:v libx264 -crf 23" and calls[0][3] == 60 * 60 * 1000
    assert c.call("XPENDING", tf.STREAM, tf.GROUP)[0] == 0
    # encoder failure: partial output removed, task ACKed and dropped, input kept
    t = mk("b")
    assert not tf.process_task(c, t, bad_exec, lambda p: None, poll=0.01)
    assert not (tmp_path / "out" / "b.mp4").exists() and os.path.exists(t.input_path)
    assert c.call("XPENDING", tf.STREAM, tf.GROUP)[0] == 0
    # verify failure: same
    t = mk("c")

    def bad_verify(p):
        raise RuntimeError("无有效视频流")
    assert not tf.process_task(c, t, ok_exec, bad_verify, poll=0.01)
    assert not (tmp_path / "out" / "c.mp4").exists()
    # missing input: times out in wait_for_file -> dropped (shortened timeout through a cancelled flag)
    t = tf.Task(id="d", input_path=str(tmp_path / "nope"), output_dir=str(tmp_path / "out"), output_name="d.mp4")
    assert not tf.process_task(c, t, ok_exec, lambda p: None, poll=0.01, cancelled=lambda: True)


def test_wait_for_file_needs_a_stable_size(tmp_path):
    p = tmp_path / "grow.bin"
    p.write_bytes(b"1")
    stop = time.time() + 0.25

    def grow():
        while time.time() < stop:
            with open(p, "ab") as f:
                f.write(b"1")
            time.sleep(0.01)
    th = threading.Thread(target=grow)
    th.start()
    t0 = time.time()
    tf.wait_for_file(str(p), timeout=5, poll=0.05)
    th.join()
    assert time.time() - t0 >= 0.25 + 2 * 0.05                   # returned only after growth stopped + 3 equal polls
    with pytest.raises(RuntimeError):
        tf.wait_for_file("", timeout=0.1, poll=0.01)


def test_consumer_pool_drops_invalid_tasks(redis, tmp_path):
    c = tf.RedisClient(*redis.addr)
    c.create_consumer_group()
    done = []
    inp = tmp_path / "in.bin"
    inp.write_bytes(b"data")
    cons = tf.Consumer(redis.addr, "gpu0", lambda i, o, a, t: (open(o, "wb").write(b"ok"), done.append(o)), lambda p: None,
                       concurrency=2, poll=0.01).start()
    c.publish(tf.Task(id="", input_path=""))                      # invalid: ACKed and dropped (consumer.go:136-142)
    c.publish(tf.Task(id="t1", input_path=str(inp), output_dir=str(tmp_path / "o"), output_name="x.mp4"))
    t0 = time.time()
    while cons.stats.processed < 1 and time.time() - t0 < 5:
        time.sleep(0.02)
    cons.shutdown(1)
    assert cons.stats.success == 1 and done and c.call("XPENDING", tf.STREAM, tf.GROUP)[0] == 0


@pytest.mark.gpu
def test_task_through_the_b200_executor(redis, tmp_path):
    """Config #1 in miniature: one task, h264-cpu preset string, --verify, through the real executor."""
    from video_codec_pipeline_b200 import api, arbiter, synth
    w, h, n = 640, 360, 12
    clip = synth.make_clip(w, h, n, seed=3)
    src = tmp_path / "in.y4m"
    with open(src, "wb") as f:
        f.write(b"YUV4MPEG2 W%d H%d F30:1 Ip A1:1 C420jpeg\n" % (w, h))
        for fr in clip:
            f.write(b"FRAME\n" + fr.tobytes())
    c = tf.RedisClient(*redis.addr)
    c.create_consumer_group()
    c.publish(tf.Task(id="t", input_path=str(src), original_name="in.y4m", output_dir=str(tmp_path / "out"),
                      output_name="in.mp4", ffmpeg_args=tf.PRESETS["h264-cpu"] + " -g 6", verify_output=True))
    cons = tf.Consumer(redis.addr, "gpu0", lambda i, o, a, t: api.transcode(i, o, a, t), api.verify, 1, poll=0.02).start()
    t0 = time.time()
    while cons.stats.processed < 1 and time.time() - t0 < 60:
        time.sleep(0.05)
    cons.shutdown(1)
    assert cons.stats.success == 1, cons.stats.log
    out = tmp_path / "out" / "in.mp4"
    assert out.exists() and not src.exists()
    if arbiter.available():
        assert len(arbiter.decode_file(str(out))) == n
