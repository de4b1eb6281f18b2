"""Reader/finaliser for the raw accumulator files ("MRTACC1") written by the
renderer front ends and by the oracle harness: a 72-byte header followed by
W*H float4 = (sum of finite radiance samples, finite-sample count)."""
import struct

import numpy as np

HEADER = struct.Struct("<8s8I2Qd")


def read_acc(path):
    with open(path, "rb") as f:
        raw = f.read(HEADER.size)
        magic, w, h, n, s0, s1, depth, scene, threads, seed, rays, secs = HEADER.unpack(raw)
        if not magic.startswith(b"MRTACC1"):
            raise ValueError(f"{path}: bad magic {magic!r}")
        acc = np.fromfile(f, dtype=np.float32, count=w * h * 4).reshape(h, w, 4)
    meta = dict(width=w, height=h, samples=n, s0=s0, s1=s1, depth=depth, scene=scene,
                threads=threads, seed=seed, rays=rays, seconds=secs)
    return acc, meta


def write_acc(path, acc, **meta):
    h, w, _ = acc.shape
    hdr = HEADER.pack(b"MRTACC1\0", w, h, meta.get("samples", 0), meta.get("s0", 0), meta.get("s1", 0),
                      meta.get("depth", 0), meta.get("scene", 0), meta.get("threads", 0),
                      meta.get("seed", 0), meta.get("rays", 0), float(meta.get("seconds", 0.0)))
    with open(path, "wb") as f:
        f.write(hdr)
        np.ascontiguousarray(acc, dtype=np.float32).tofile(f)


def finalize(acc, max_luminance=1000.0):
    """mean over finite samples, then the luminance clamp of main.cpp:168-173 (float32 throughout)."""
    acc = np.asarray(acc, dtype=np.float32)
    cnt = acc[..., 3:4]
    with np.errstate(divide="ignore", invalid="ignore"):
        color = np.where(cnt > 0, acc[..., :3] / cnt, np.float32(0)).astype(np.float32)
    c = np.array([0.212655, 0.715158, 0.072187], dtype=np.float32)
    prod = color * c
    lum = (prod[..., 0] + prod[..., 1]) + prod[..., 2]
    over = lum > np.float32(max_luminance)
    with np.errstate(divide="ignore", invalid="ignore"):
        scale = np.where(over, np.float32(max_luminance) / lum, np.float32(1)).astype(np.float32)
    return np.where(over[..., None], color * scale[..., None], color).astype(np.float32)


def compare(a, b, rel=1e-4, abs_floor=1e-6):
    """Per-pixel relative agreement of two finalised images: a pixel agrees when every channel is
    within rel * max(|a|, |b|) (+ a tiny absolute floor for black pixels)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    tol = rel * np.maximum(np.abs(a), np.abs(b)) + abs_floor
    ok = np.all(np.abs(a - b) <= tol, axis=-1)
    diff = np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), abs_floor)
    return dict(frac_ok=float(ok.mean()), n_bad=int((~ok).sum()), max_rel=float(diff.max()),
                rmse=float(np.sqrt(np.mean((a - b) ** 2))))
