"""Sliding-triplet driver loop and frame-window sharding (host logic, device agnostic).

Mirrors the reference's driver (infer.py:58-174): timestamp schedule ``calc_t`` (infer.py:76-91,
models/utils/tools.py:120-134), head / loop / tail, and the 4-way scene-cut state machine
(infer.py:122-143), including its quirks (``idx`` is not advanced by the
head, so head and first window both use ``calc_t(0)``; ``reuse`` is reset on every scene cut).

Multi-GPU: the stream is partitioned into contiguous ranges of loop iterations,
one replica per GPU, no exchange step.  A shard that starts at iteration ``a > 0`` rebuilds the
``reuse`` state exactly as the sequential loop would have left it (one ``calc_flow`` on frames
a, a+1), so the concatenation of the shards' outputs equals the sequential run's output.
"""
import math

import numpy as np


class TMapper:
    """models/utils/tools.py:120-134."""

    def __init__(self, src=-1., dst=0., times=-1):
        self.times = dst / src if times == -1 else times
        self.now_step = -1

    def get_range_timestamps(self, _min, _max, lclose=True, rclose=False, normalize=True):
        _min_step = math.ceil(_min * self.times)
        _max_step = math.ceil(_max * self.times)
        _start = _min_step if lclose else _min_step + 1
        _end = _max_step if not rclose else _max_step + 1
        if _start >= _end:
            return []
        if normalize:
            return [((_i / self.times) - _min) / (_max - _min) for _i in range(_start, _end)]
        return [_i / self.times for _i in range(_start, _end)]


def make_calc_t(src_fps, dst_fps, times=-1):
    """infer.py:73-91: timestamps in [0.5, 1.5) relative to the window's first frame."""
    t_mapper = TMapper(src_fps, dst_fps, times)

    def calc_t(_idx):
        if times != -1:
            if times % 2:
                vfi_timestamp = [(_i + 1) / times for _i in range((times - 1) // 2)]
                vfi_timestamp = list(reversed([1 - t for t in vfi_timestamp])) + [1] + [t + 1 for t in vfi_timestamp]
                return np.array(vfi_timestamp)
            vfi_timestamp = [(_i + 0.5) / times for _i in range(times // 2)]
            vfi_timestamp = list(reversed([1 - t for t in vfi_timestamp])) + [t + 1 for t in vfi_timestamp]
            return np.array(vfi_timestamp)
        timestamp = np.array(t_mapper.get_range_timestamps(_idx - 0.5, _idx + 0.5, lclose=True, rclose=False,
                                                           normalize=False))
        return np.round(timestamp - _idx, 4) + 1

    return calc_t


def _no_scene(_a, _b):
    return False


def head_outputs(model, I0, I1, ts, scene):
    """infer.py:93-107."""
    if scene:
        return [I0 for _ in ts]
    left_ts = ts[ts < 1]
    right_ts = ts[ts >= 1] - 1
    output = [I0 for _ in left_ts]
    output.extend(model.inference_ts(I0, I1, right_ts))
    return output


def window_outputs(model, I0, I1, I2, ts, reuse, left_scene, right_scene):
    """One loop iteration of infer.py:112-143; returns (outputs, reuse)."""
    if left_scene and right_scene:
        return [I1 for _ in ts], None
    if left_scene and not right_scene:
        left_ts = ts[ts < 1]
        right_ts = ts[ts >= 1] - 1
        output = [I1 for _ in left_ts]
        output.extend(model.inference_ts(I1, I2, right_ts))
        return output, None
    if not left_scene and right_scene:
        left_ts = ts[ts <= 1]
        right_ts = ts[ts > 1] - 1
        output = model.inference_ts(I0, I1, left_ts)
        output.extend([I1 for _ in right_ts])
        return output, None
    return model.inference_ts_drba(I0, I1, I2, ts, reuse, linear=True)


def tail_outputs(model, I0, I1, ts):
    """infer.py:158-164."""
    left_ts = ts[ts <= 1]
    right_ts = ts[ts > 1] - 1
    output = model.inference_ts(I0, I1, left_ts)
    output.extend([I1 for _ in right_ts])
    return output


def interpolate_sequence(model, frames, src_fps, dst_fps, times=-1, check_scene=None):
    """The whole reference loop over an indexable sequence of network-ready frames.
    Yields output frames in order.  ``check_scene(a, b) -> bool`` or None (scene detection off)."""
    yield from interpolate_shard(model, frames, src_fps, dst_fps, times, check_scene, 0, None)


def num_iterations(n_frames):
    """Loop iterations of the reference for an n-frame clip (one per frame after the first two)."""
    return max(n_frames - 2, 0)


def shard_ranges(n_iterations, world):
    """Contiguous, balanced ranges [a, b) of loop iterations, one per rank.  Ranks without work (world >
    n_iterations) get (None, None): interpolate_shard emits nothing for them, so head and tail are emitted exactly
    once (by the shard that starts at 0 and the last non-empty one; rank 0 when the clip has no iteration at all)."""
    base, rem = divmod(n_iterations, world)
    out, a = [], 0
    for r in range(world):
        b = a + base + (1 if r < rem else 0)
        out.append((a, b) if (b > a or (n_iterations == 0 and r == 0)) else (None, None))
        a = b
    return out


def shard_reuse(model, Ia, Ib):
    """The `reuse` the sequential loop hands to iteration a, rebuilt from frames a, a+1.
    RIFE (models/rife.py:109): (flow21, flow12, f2, f1) of the previous window = calc_flow(Ia, Ib) pair-swapped;
    GMFSS / GMFSS_UNION (models/gmfss.py:71, gmfss_union.py:98): Model.reuse(Ia, Ib) pair-swapped."""
    if hasattr(model, "shard_reuse"):
        return model.shard_reuse(Ia, Ib)
    f01, f10, fa, fb = model.calc_flow(Ia, Ib)
    return (f10, f01, fb, fa)


def interpolate_shard(model, frames, src_fps, dst_fps, times=-1, check_scene=None, a=0, b=None):
    """Outputs of loop iterations [a, b) (b=None: to the end of the clip).  The shard containing
    iteration 0 also emits the head; the shard containing the last iteration also emits the tail.
    Iteration j uses frames j, j+1, j+2 and ``calc_t(j)``."""
    n = len(frames)
    if n < 2:
        raise ValueError("need at least two frames")
    n_it = num_iterations(n)
    b = n_it if b is None else b
    calc_t = make_calc_t(src_fps, dst_fps, times)
    scene = check_scene if check_scene is not None else _no_scene
    if a is None:
        return          # a rank without work (shard_ranges: world > n_iterations)
    a, b = max(0, min(a, n_it)), max(0, min(b, n_it))
    if a == b and n_it > 0:
        return          # an empty range of a non-empty clip emits nothing (head / tail belong to non-empty shards)
    first, last = a == 0, b == n_it

    if first:
        left_scene = bool(scene(frames[0], frames[1]))
        for x in head_outputs(model, frames[0], frames[1], calc_t(0), left_scene):
            yield x
        reuse = None
    elif a < b:
        # state the sequential loop holds when it enters iteration a
        left_scene = bool(scene(frames[a], frames[a + 1]))
        prev_left = bool(scene(frames[a - 1], frames[a]))
        if not prev_left and not left_scene:
            reuse = shard_reuse(model, frames[a], frames[a + 1])
        else:
            reuse = None
    for j in range(a, b):
        right_scene = bool(scene(frames[j + 1], frames[j + 2]))
        out, reuse = window_outputs(model, frames[j], frames[j + 1], frames[j + 2], calc_t(j), reuse,
                                    left_scene, right_scene)
        for x in out:
            yield x
        left_scene = right_scene
    if last:
        # tail: the loop leaves I0, I1 = frames[n-2], frames[n-1] and idx = n_it
        for x in tail_outputs(model, frames[n - 2], frames[n - 1], calc_t(n_it)):
            yield x
