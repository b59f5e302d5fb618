#!/usr/bin/env python
"""DRBA command line, re-hosted on the B200 implementation of the hot path.

Same flags and loop semantics as the reference's infer.py (:18-36 flags, :58-174 loop); the
per-window work runs in libdrba_b200.so.  `-m rife`, `-m gmfss` and `-m gmfss_union` are served; other model
names raise like the reference's unknown-model branch.
Extra flags: --precision {fp16,fp32} (conv engine), --weights DIR, --gpus N (frame-window sharding over N GPUs).
"""
import argparse
import os
import subprocess
import shutil
import sys
import time
from queue import Queue
from threading import Thread

import numpy as np


def parse_args():
    parser = argparse.ArgumentParser(description='Interpolation a video with DRBA')
    parser.add_argument('-m', '--model_type', dest='model_type', type=str, default='rife',
                        help='model network type, current support rife/gmfss/gmfss_union')
    parser.add_argument('-i', '--input', dest='input', type=str, default='input.mp4', help='absolute path of input video')
    parser.add_argument('-o', '--output', dest='output', type=str, default='output.mp4', help='absolute path of output video')
    parser.add_argument('-fps', '--dst_fps', dest='dst_fps', type=float, default=60, help='interpolate to ? fps')
    parser.add_argument('-t', '--times', dest='times', type=int, default=-1, help='interpolate to ?x fps')
    parser.add_argument('-s', '--enable_scdet', dest='enable_scdet', action='store_true', default=False,
                        help='enable scene change detection')
    parser.add_argument('-st', '--scdet_threshold', dest='scdet_threshold', type=float, default=0.3,
                        help='ssim scene detection threshold')
    parser.add_argument('-hw', '--hwaccel', dest='hwaccel', action='store_true', default=False,
                        help='enable hardware acceleration encode(require nvidia graph card)')
    parser.add_argument('-scale', '--scale', dest='scale', type=float, default=1.0,
                        help='flow scale, generally use 1.0 with 1080P and 0.5 with 4K resolution')
    parser.add_argument('--precision', default='fp16', choices=['fp16', 'fp32'])
    parser.add_argument('--weights', default=None, help='directory holding flownet.pkl')
    parser.add_argument('--gpus', type=int, default=1,
                        help='shard the frame stream over this many GPUs (one replica process each, no collective)')
    parser.add_argument('--shard-buffer', dest='shard_buffer', type=int, default=256,
                        help='output frames a shard may run ahead of the writer (--gpus > 1)')
    return parser.parse_args()


def load_model(model_type, scale, device, precision, weights):
    if model_type == 'rife':
        from drba_b200.rife import RIFE
        from drba_b200.weights import find_rife_weights
        wdir = find_rife_weights(weights)
        if wdir is None:
            raise FileNotFoundError('weights/train_log_rife_426_heavy/flownet.pkl')
        return RIFE(weights=wdir, scale=scale, device=device, precision=precision)
    if model_type == 'gmfss':
        from drba_b200.gmfss import GMFSS
        from drba_b200.weights import find_gmfss_weights
        wdir = find_gmfss_weights(weights)
        if wdir is None:
            raise FileNotFoundError('weights/train_log_gmfss/fusionnet.pkl')
        return GMFSS(weights=wdir, scale=scale, device=device)
    if model_type == 'gmfss_union':
        from drba_b200.gmfss_union import GMFSS_UNION
        wdir = weights or 'weights/train_log_gmfss_union'
        for cand in (wdir, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'baseline/_ref/weights/train_log_gmfss_union'),
                     '/root/reference/weights/train_log_gmfss_union'):
            if os.path.isfile(os.path.join(cand, 'fusionnet.pkl')):
                return GMFSS_UNION(weights=cand, scale=scale, device=device, precision=precision)
        raise FileNotFoundError('weights/train_log_gmfss_union/fusionnet.pkl')
    raise ValueError(f'model_type must in {model_type}')


class VideoIO:
    """models/utils/tools.py:156-213: cv2 decode thread + encoder thread.  The encoder is the
    reference's ffmpeg rawvideo pipe when an ffmpeg binary exists, else cv2.VideoWriter."""

    def __init__(self, input_path, output_path, dst_fps=60, times=-1, hwaccel=False, read=True):
        import cv2
        self.cv2 = cv2
        self.cap = cv2.VideoCapture(input_path)
        self.src_fps = self.cap.get(cv2.CAP_PROP_FPS)
        self.dst_fps = dst_fps if times == -1 else times * self.src_fps
        self.total_frames_count = self.cap.get(7)
        self.width = int(self.cap.get(cv2.CAP_PROP_FRAME_WIDTH))
        self.height = int(self.cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
        self.ffmpeg = None
        self.writer = None
        if shutil.which('ffmpeg'):
            enc, preset = ('h264_nvenc', 'p7') if hwaccel else ('libx264', 'medium')
            cmd = ['ffmpeg', '-y', '-f', 'rawvideo', '-pix_fmt', 'rgb24', '-r', f'{self.dst_fps}',
                   '-s', f'{self.width}x{self.height}', '-i', 'pipe:0', '-i', input_path, '-map', '0:v', '-map', '1:a?',
                   '-c:v', enc, '-movflags', '+faststart', '-pix_fmt', 'yuv420p', '-qp', '16', '-preset', preset,
                   '-c:a', 'aac', '-b:a', '320k', f'{output_path}']
            self.ffmpeg = subprocess.Popen(cmd, stdin=subprocess.PIPE)
        else:
            self.writer = cv2.VideoWriter(output_path, cv2.VideoWriter_fourcc(*'mp4v'), self.dst_fps, (self.width, self.height))
        self.read_buffer = Queue(maxsize=100)
        self.write_buffer = Queue(maxsize=-1)
        self.done = False
        self.error = None
        if read:
            Thread(target=self._read, daemon=True).start()
        self._writer_thread = Thread(target=self._write, daemon=True)
        self._writer_thread.start()

    def _read(self):
        ret, x = self.cap.read()
        while ret:
            self.read_buffer.put(x)
            ret, x = self.cap.read()
        self.read_buffer.put(None)

    def _write(self):
        try:
            while True:
                item = self.write_buffer.get()
                if item is None:
                    break
                if self.ffmpeg is not None:
                    self.ffmpeg.stdin.write(np.ascontiguousarray(item[:, :, ::-1]))   # BGR -> RGB only here (tools.py:202)
                else:
                    self.writer.write(item)
            if self.ffmpeg is not None:
                self.ffmpeg.stdin.close()
                self.ffmpeg.wait()
            else:
                self.writer.release()
        except BaseException as e:      # encoder died (bad codec, -hw without nvenc): finish() re-raises instead of hanging
            self.error = e
        finally:
            self.done = True

    def write_frame(self, x):
        if self.error is not None:
            raise RuntimeError(f"video encoder failed: {self.error!r}") from self.error
        self.write_buffer.put(x)

    def read_frame(self):
        return self.read_buffer.get()

    def finish(self):
        self.write_buffer.put(None)
        self._writer_thread.join()
        if self.error is not None:
            raise RuntimeError(f"video encoder failed: {self.error!r}") from self.error
        if self.ffmpeg is not None and self.ffmpeg.returncode not in (0, None):
            raise RuntimeError(f"ffmpeg exited with code {self.ffmpeg.returncode}")


def inference(args):
    import torch
    from drba_b200 import driver, tools
    device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')
    model = load_model(args.model_type, args.scale, device, args.precision, args.weights)
    io = VideoIO(args.input, args.output, dst_fps=args.dst_fps, times=args.times, hwaccel=args.hwaccel)
    if io.dst_fps <= io.src_fps:
        raise ValueError(f'dst fps should be greater than src fps, but got dst_fps={io.dst_fps} and src_fps={io.src_fps}')
    try:
        from tqdm import tqdm
        pbar = tqdm(total=io.total_frames_count)
    except Exception:
        pbar = None
    i0, i1 = io.read_frame(), io.read_frame()
    size = tools.get_valid_net_inp_size(i0, model.scale, div=model.pad_size)
    src_size, dst_size = size['src_size'], size['dst_size']
    calc_t = driver.make_calc_t(io.src_fps, io.dst_fps, args.times)
    # scene detection: one kernel per frame pair writing its flag to pinned host memory (tools.SceneDetector); the
    # pair (I1, I2) is submitted as soon as I2 is uploaded -- one frame AHEAD of the window that needs it -- so the
    # flag has landed when the loop reads it and the stream never drains (the reference syncs three times per pair)
    det = tools.SceneDetector(device, args.scdet_threshold) if args.enable_scdet else None

    def submit(a, b):
        return det.submit(a, b) if det is not None else None

    def flag(ticket):
        return det.result(ticket) if det is not None else False

    def emit(frames):
        for x in frames:
            io.write_frame(tools.to_out(x, src_size))
        if pbar is not None:
            pbar.update(1)

    I0, I1 = tools.to_inp(i0, dst_size, device), tools.to_inp(i1, dst_size, device)
    idx = 0
    left_scene = flag(submit(I0, I1))
    reuse = None
    i2 = io.read_frame()
    I2 = tools.to_inp(i2, dst_size, device) if i2 is not None else None
    right_ticket = submit(I1, I2) if I2 is not None else None
    emit(driver.head_outputs(model, I0, I1, calc_t(idx), left_scene))
    while I2 is not None:
        i3 = io.read_frame()                     # look-ahead: upload and scene-check the NEXT pair before this window runs
        I3 = tools.to_inp(i3, dst_size, device) if i3 is not None else None
        next_ticket = submit(I2, I3) if I3 is not None else None
        right_scene = flag(right_ticket)
        output, reuse = driver.window_outputs(model, I0, I1, I2, calc_t(idx), reuse, left_scene, right_scene)
        emit(output)
        I0, I1, I2 = I1, I2, I3
        left_scene, right_ticket = right_scene, next_ticket
        idx += 1
    emit(driver.tail_outputs(model, I0, I1, calc_t(idx)))
    io.finish()
    if pbar is not None:
        pbar.close()


# ---- frame-window sharding over several GPUs (BASELINE.json configs[4]) ----------------------------
class ClipFrames:
    """Indexable, forward-moving view of a video's network-ready frames for driver.interpolate_shard: decodes on
    demand, keeps the last few frames on the device.  Frames before the first requested index are skipped with
    cap.grab() (no colour conversion, no upload)."""

    def __init__(self, path, n_frames, dst_size, device, keep=6):
        import cv2
        from drba_b200 import tools
        self.cap, self.n, self.dst, self.device, self.keep = cv2.VideoCapture(path), n_frames, dst_size, device, keep
        self.tools = tools
        self.pos = 0               # index of the next frame the decoder returns
        self.cache = {}

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        if i < 0:
            i += self.n
        if i in self.cache:
            return self.cache[i]
        if i < self.pos:           # the driver only moves forward, except for the head of shard 0
            import cv2
            self.cap.set(cv2.CAP_PROP_POS_FRAMES, 0)
            self.pos = 0
        while self.pos < i:
            self.cap.grab()
            self.pos += 1
        ok, x = self.cap.read()
        if not ok:
            raise IndexError(i)
        self.pos += 1
        self.cache[i] = self.tools.to_inp(x, self.dst, self.device)
        for k in sorted(self.cache)[:-self.keep]:
            del self.cache[k]
        return self.cache[i]


def count_frames(path):
    import cv2
    cap = cv2.VideoCapture(path)
    n = 0
    while cap.grab():
        n += 1
    return n


def _shard_worker(rank, world, args, a, b, n_frames, queue):
    """One replica: iterations [a, b) of the reference loop on GPU `rank`, output frames to the parent in order."""
    try:
        import torch
        from drba_b200 import driver, tools
        ndev = torch.cuda.device_count()
        device = torch.device('cuda', rank % max(ndev, 1))
        torch.cuda.set_device(device)
        model = load_model(args.model_type, args.scale, device, args.precision, args.weights)
        import cv2
        cap = cv2.VideoCapture(args.input)
        src_fps = cap.get(cv2.CAP_PROP_FPS)
        ok, first = cap.read()
        dst_fps = args.dst_fps if args.times == -1 else args.times * src_fps
        size = tools.get_valid_net_inp_size(first, model.scale, div=model.pad_size)
        frames = ClipFrames(args.input, n_frames, size['dst_size'], device)
        if a > 0:
            frames[a - 1]          # the state rebuild looks one frame back: decode it before moving on
        scene = (lambda x, y: tools.check_scene(x, y, args.scdet_threshold)) if args.enable_scdet else None
        for x in driver.interpolate_shard(model, frames, src_fps, dst_fps, args.times, scene, a, b):
            queue.put(tools.to_out(x, size['src_size']))
        queue.put(None)
    except BaseException as e:      # the parent re-raises
        import traceback
        queue.put(RuntimeError(f"shard {rank} failed: {e!r}\n{traceback.format_exc()}"))


def inference_sharded(args):
    """`--gpus N`: the frame stream is cut into N contiguous ranges of loop iterations, one replica process per GPU,
    no exchange step; a shard that starts mid-stream rebuilds `reuse` exactly as the sequential loop leaves it
    (drba_b200.driver.interpolate_shard), so the written clip equals the single-GPU run's."""
    import multiprocessing as mp
    from drba_b200 import driver
    n_frames = count_frames(args.input)
    if n_frames < 2:
        raise ValueError('need at least two frames')
    io = VideoIO(args.input, args.output, dst_fps=args.dst_fps, times=args.times, hwaccel=args.hwaccel, read=False)
    if io.dst_fps <= io.src_fps:
        raise ValueError(f'dst fps should be greater than src fps, but got dst_fps={io.dst_fps} and src_fps={io.src_fps}')
    ranges = driver.shard_ranges(driver.num_iterations(n_frames), args.gpus)
    ctx = mp.get_context('spawn')
    queues = [ctx.Queue(maxsize=args.shard_buffer) for _ in ranges]
    procs = []
    for r, (a, b) in enumerate(ranges):
        if a is None:
            procs.append(None)
            continue
        p = ctx.Process(target=_shard_worker, args=(r, len(ranges), args, a, b, n_frames, queues[r]), daemon=True)
        p.start()
        procs.append(p)
    for r, p in enumerate(procs):          # the writer concatenates the shards' outputs in shard order
        if p is None:
            continue
        while True:
            item = queues[r].get()
            if item is None:
                break
            if isinstance(item, BaseException):
                raise item
            io.write_frame(item)
        p.join()
    io.finish()


if __name__ == '__main__':
    a = parse_args()
    if not os.path.exists(a.input):
        raise FileNotFoundError(f"can't find the video file {a.input}")
    if a.gpus > 1:
        inference_sharded(a)
    else:
        inference(a)
