"""ctypes front-end of oracle/cpu_hough_ransac.cpp — the reference's CPU `Houghvoting` op
(lib/hough_voting_layer, pre-emptive RANSAC) restated in standalone C++.

TEST INFRASTRUCTURE / CPU BASELINE ONLY: used by tests, by bench.py's cpu_baseline leg and by
`bench.py --impl reference`.  Never imported by the product package.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "_build", "libcpu_hough.so")
        src = os.path.join(_HERE, "cpu_hough_ransac.cpp")
        if not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(so) < os.path.getmtime(src)):
            subprocess.check_call(["make", "-C", _HERE, "_build/libcpu_hough.so"], stdout=subprocess.DEVNULL)
        _LIB = ctypes.CDLL(so)
    return _LIB


def hough_voting(label, vertex, extents, meta, is_train=0, threads=1):
    """Houghvoting (CPU op, hough_voting_op.cc:38-49): returns top_box [R,6], top_pose [R,7]."""
    label = np.ascontiguousarray(label, np.int32); vertex = np.ascontiguousarray(vertex, np.float32)
    extents = np.ascontiguousarray(extents, np.float32); meta = np.ascontiguousarray(meta, np.float32)
    B, H, W = label.shape
    C = vertex.shape[3] // 3
    cap = B * C * 9 * 256 + 1  # train mode keeps every hypothesis that survived the halvings (getWorkingQueue, :372-377)
    box = np.zeros((cap, 6), np.float32); pose = np.zeros((cap, 7), np.float32)
    fp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    n = lib().cpu_hough_voting(label.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), fp(vertex), fp(extents), fp(meta), B, H, W,
                               C, meta.shape[-1], int(is_train), int(threads), fp(box), fp(pose), cap)
    assert n >= 1
    return box[:n].copy(), pose[:n].copy()


def _bench_frames(nframes, seed0=2234):
    from posecnn_b200 import synth
    return synth.make_scene(batch=nframes, height=480, width=640, num_classes=22, seed=seed0)


def timed_baseline(sample_frames=2, threads=None, repeats=3):
    """frames/s of the CPU op on `sample_frames` synthetic 640x480 / 22-class frames (same generator and
    shapes as the GPU workload); best-effort variant: all host cores (OpenMP on), -O3."""
    cores = threads or os.cpu_count() or 1
    sc = _bench_frames(sample_frames)
    args = (sc["label"], sc["vertex"], sc["extents"], sc["meta"])
    hough_voting(*args, threads=cores)  # warm-up
    ts = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        hough_voting(*args, threads=cores)
        ts.append(time.perf_counter() - t0)
    t1 = []
    t0 = time.perf_counter()
    hough_voting(*args, threads=1)
    t1 = time.perf_counter() - t0
    dt = float(np.median(ts))
    return dict(value=sample_frames / dt, unit="frames/s", cores=cores, kind="port",
                sample=f"{sample_frames} synthetic 640x480x22-class frames, CPU hough_voting_layer (RANSAC) restated in C++ "
                       f"(-O3 -fopenmp, {cores} threads; as-shipped 1-thread: {sample_frames / t1:.2f} frames/s), "
                       f"median of {repeats}", ms_per_frame=1e3 * dt / sample_frames,
                single_thread_frames_per_s=sample_frames / t1)


def reference_arm(args):
    """`bench.py --impl reference`: the reference's own CPU implementation of the path on host cores."""
    cores = os.cpu_count() or 1
    frames = 2
    sc = _bench_frames(frames)
    a = (sc["label"], sc["vertex"], sc["extents"], sc["meta"])
    for _ in range(max(1, min(args.warmup, 2))):
        hough_voting(*a, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        hough_voting(*a, threads=cores)
    dt = (time.perf_counter() - t0) / args.steps
    v = frames / dt
    return dict(impl="reference", metric="frames/sec on 640x480, 21 classes (Hough voting op)", value=v, unit="frames/s",
                n_gpus=0, steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * dt, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f64/f32", data="synthetic",
                config=dict(workload=f"CPU hough_voting_layer (RANSAC), {frames} frames 640x480x22 per step"),
                cpu_baseline=dict(value=v, unit="frames/s", cores=cores, kind="port",
                                  sample=f"{frames} frames per step, {args.steps} steps, {cores} OpenMP threads"),
                e2e=dict(value=v, unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
