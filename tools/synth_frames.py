"""Deterministic synthetic 10-bit luma frames with known global affine motion.

SURVEY.md 8(d): multi-octave value noise (lattices 64/16/4 px, amplitudes
256/128/64 around 512) sampled through a per-frame global affine map
(zoom 1+0.0015 t, rotation 0.05 deg * t about the centre, translation
(0.75 t, -0.5 t) px).  "Original" sequence = frames 1..N, "reconstructed"
(reference) sequence = frames 0..N-1 plus uniform noise in [-a, a] with
a = {22:1, 27:2, 32:3, 37:5}[QP].

Also writes / reads the reference's CSV frame format (main.cpp:303-328): H lines
of W comma-separated decimal samples per frame, frames stacked vertically.
"""
import numpy as np

SEED = 0xA11F1E5
_NOISE_AMP = {22: 1, 27: 2, 32: 3, 37: 5}


def _hash01(ix, iy, salt):
    """Integer lattice hash -> float in [-1, 1) (SplitMix64 finaliser)."""
    with np.errstate(over="ignore"):
        z = (ix.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
             + iy.astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F) + np.uint64(salt))
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 / (1 << 53)) - 1.0


def _value_noise(x, y, cell, salt):
    gx, gy = x / cell, y / cell
    x0, y0 = np.floor(gx), np.floor(gy)
    fx, fy = gx - x0, gy - y0
    fx = fx * fx * (3 - 2 * fx)
    fy = fy * fy * (3 - 2 * fy)
    ix, iy = x0.astype(np.int64) + (1 << 20), y0.astype(np.int64) + (1 << 20)
    v00 = _hash01(ix, iy, salt)
    v10 = _hash01(ix + 1, iy, salt)
    v01 = _hash01(ix, iy + 1, salt)
    v11 = _hash01(ix + 1, iy + 1, salt)
    return (v00 * (1 - fx) + v10 * fx) * (1 - fy) + (v01 * (1 - fx) + v11 * fx) * fy


def frame(t, W, H, seed=SEED):
    """Frame t of the sequence as (H, W) uint16 in [0, 1023]."""
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float64)
    cx, cy = (W - 1) / 2.0, (H - 1) / 2.0
    zoom = 1.0 + 0.0015 * t
    ang = np.deg2rad(0.05 * t)
    ca, sa = np.cos(ang) / zoom, np.sin(ang) / zoom
    dx, dy = xs - cx, ys - cy
    u = ca * dx + sa * dy + cx - 0.75 * t
    v = -sa * dx + ca * dy + cy + 0.5 * t
    val = 512.0
    for cell, amp, k in ((64.0, 256.0, 1), (16.0, 128.0, 2), (4.0, 64.0, 3)):
        val = val + amp * _value_noise(u, v, cell, seed * 16 + k)
    return np.clip(np.rint(val), 0, 1023).astype(np.uint16)


def sequences(n_frames, W, H, qp=32, seed=SEED):
    """Returns (original[n,H,W], reconstructed[n,H,W]) uint16: original = frames
    1..n, reconstructed = frames 0..n-1 plus coding-noise."""
    frames = [frame(t, W, H, seed) for t in range(n_frames + 1)]
    orig = np.stack(frames[1:])
    a = _NOISE_AMP.get(qp, 3)
    rng = np.random.Generator(np.random.PCG64(seed + 1000 * qp))
    noise = rng.integers(-a, a + 1, size=(n_frames, H, W))
    recon = np.clip(np.stack(frames[:-1]).astype(np.int64) + noise, 0, 1023).astype(np.uint16)
    return orig, recon


def stress_frames(W, H, seed=SEED):
    """Parity-only stress planes: (name, cur, ref) triples."""
    rng = np.random.Generator(np.random.PCG64(seed + 7))
    xs = np.arange(W, dtype=np.int64)[None, :].repeat(H, 0)
    ys = np.arange(H, dtype=np.int64)[:, None].repeat(W, 1)
    const = np.full((H, W), 512, np.uint16)
    ramp = np.clip(xs * 1023 // max(W - 1, 1), 0, 1023).astype(np.uint16)
    ramp2 = np.clip((xs + 3) * 1023 // max(W - 1, 1), 0, 1023).astype(np.uint16)
    edge = np.where(xs < W // 2 + 5, 100, 900).astype(np.uint16)
    edge2 = np.where(xs < W // 2 + 7, 100, 900).astype(np.uint16)
    chk = np.where(((xs // 8) + (ys // 8)) % 2 == 0, 200, 800).astype(np.uint16)
    chk2 = np.where((((xs + 2) // 8) + ((ys + 1) // 8)) % 2 == 0, 200, 800).astype(np.uint16)
    wn1 = rng.integers(0, 1024, size=(H, W)).astype(np.uint16)
    wn2 = rng.integers(0, 1024, size=(H, W)).astype(np.uint16)
    zeros = np.zeros((H, W), np.uint16)
    full = np.full((H, W), 1023, np.uint16)
    return [("const", const, const), ("ramp", ramp, ramp2), ("edge", edge, edge2), ("checker", chk, chk2),
            ("noise", wn1, wn2), ("zeros", zeros, zeros), ("full_vs_zero", full, zeros)]


def write_csv(path, planes):
    """planes: (n, H, W) -> the reference's CSV layout."""
    planes = np.asarray(planes)
    n, H, W = planes.shape
    with open(path, "w") as f:
        for k in range(n):
            np.savetxt(f, planes[k], fmt="%d", delimiter=",")


def read_csv(path, n, W, H):
    data = np.loadtxt(path, delimiter=",", dtype=np.int64, max_rows=n * H)
    return data.reshape(n, H, W).astype(np.uint16)
