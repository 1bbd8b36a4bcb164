"""Deterministic synthetic frame sources: integer only, no files.

noise(seed)  every byte from the 64-bit LCG  s = s*6364136223846793005 + 1442695040888963407,
             byte = s >> 56
camera(t)    moving gradients + four moving textured squares (I420, or YUYV for config 1)
screen(t)    static black-on-white glyph cells with a scrolling band (screen-share content)
sports(t)    fast pan (37 samples per picture), a fast foreground object and a scene cut at t = 3
"""
from __future__ import annotations

import numpy as np

_A = np.uint64(6364136223846793005)
_C = np.uint64(1442695040888963407)


def lcg_bytes(seed: int, n: int) -> np.ndarray:
    """n bytes of the LCG stream that starts from state `seed` (first byte uses the first update)."""
    if n <= 0:
        return np.zeros(0, np.uint8)
    with np.errstate(over="ignore"):
        a_pow = np.cumprod(np.full(n, _A, dtype=np.uint64))            # a^1 .. a^n
        geo = np.empty(n, dtype=np.uint64)                             # 1 + a + .. + a^(k-1)
        geo[0] = 1
        if n > 1:
            geo[1:] = np.cumsum(a_pow[:-1], dtype=np.uint64) + np.uint64(1)
        s = a_pow * np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + _C * geo
    return (s >> np.uint64(56)).astype(np.uint8)


def noise(seed: int, nbytes: int) -> np.ndarray:
    return lcg_bytes(seed, nbytes)


def i420_size(w: int, h: int) -> int:
    return w * h * 3 // 2


def camera_i420(w: int, h: int, t: int) -> np.ndarray:
    """Packed I420 frame t of the `camera` sequence."""
    x = np.arange(w, dtype=np.int64)[None, :]
    y = np.arange(h, dtype=np.int64)[:, None]
    Y = 16 + ((3 * x + 2 * y + 5 * t) % 220)
    tex = lcg_bytes(77, 64 * 64).reshape(64, 64).astype(np.int64)
    for k, (vx, vy) in enumerate(((3, 2), (-3, 2), (3, -2), (-3, -2))):
        px = (w // 5 * (k + 1) + vx * t) % max(w - 64, 1)
        py = (h // 5 * (k + 1) + vy * t) % max(h - 64, 1)
        hh = min(64, h - py)
        ww = min(64, w - px)
        blk = Y[py:py + hh, px:px + ww]
        Y[py:py + hh, px:px + ww] = (blk + tex[:hh, :ww]) // 2
    cx = np.arange(w // 2, dtype=np.int64)[None, :] * 2
    cy = np.arange(h // 2, dtype=np.int64)[:, None] * 2
    U = 128 + (((cx >> 3) + t) % 32) - 16 + 0 * cy
    V = 128 + (((cy >> 3) - t) % 32) - 16 + 0 * cx
    out = np.empty(i420_size(w, h), np.uint8)
    out[: w * h] = np.clip(Y, 0, 255).astype(np.uint8).ravel()
    out[w * h: w * h + w * h // 4] = U.astype(np.uint8).ravel()
    out[w * h + w * h // 4:] = V.astype(np.uint8).ravel()
    return out


def i420_to_yuyv(i420: np.ndarray, w: int, h: int) -> np.ndarray:
    """Packed YUY2 with chroma rows duplicated (a camera that upsampled 4:2:0 by repetition)."""
    Y = i420[: w * h].reshape(h, w)
    U = i420[w * h: w * h + w * h // 4].reshape(h // 2, w // 2)
    V = i420[w * h + w * h // 4:].reshape(h // 2, w // 2)
    out = np.empty((h, w // 2, 4), np.uint8)
    out[:, :, 0] = Y[:, 0::2]
    out[:, :, 2] = Y[:, 1::2]
    out[:, :, 1] = np.repeat(U, 2, axis=0)
    out[:, :, 3] = np.repeat(V, 2, axis=0)
    return out.ravel()


def screen_i420(w: int, h: int, t: int) -> np.ndarray:
    """Screen-share content: 8x16 glyph cells, black on white, rows 20..40 scroll every 15 frames."""
    cols, rows = (w + 7) // 8, (h + 15) // 16
    Y = np.full((rows * 16, cols * 8), 255, np.uint8)
    shift = t // 15
    for r in range(rows):
        src_r = r + shift if 20 <= r <= 40 else r
        bits = lcg_bytes(1000 + src_r, cols * 16)              # one byte = one 8-pixel glyph row
        g = np.unpackbits(bits.reshape(cols, 16, 1), axis=2)   # cols,16,8
        blockrow = np.where(g.transpose(1, 0, 2).reshape(16, cols * 8) == 1, 0, 255).astype(np.uint8)
        # thin the glyphs so that text stays legible-like: keep ~1/3 of the ink
        mask = (lcg_bytes(5000 + src_r, cols * 16 * 8).reshape(16, cols * 8) % 3) == 0
        Y[r * 16:(r + 1) * 16] = np.where(mask, blockrow, 255)
    out = np.full(i420_size(w, h), 128, np.uint8)
    out[: w * h] = Y[:h, :w].ravel()
    return out


def _world(seed: int, w: int, h: int) -> np.ndarray:
    """Textured plane: 8x8-blocky random base (strong structure a block matcher can lock onto) plus
    fine-grained detail, integer only."""
    bw, bh = (w + 7) // 8 + 1, (h + 7) // 8 + 1
    base = lcg_bytes(seed, bw * bh).reshape(bh, bw).astype(np.int64)
    big = np.kron(base, np.ones((8, 8), np.int64))[:h, :w]
    fine = lcg_bytes(seed + 1, w * h).reshape(h, w).astype(np.int64)
    return 32 + (3 * big + fine) // 5


def sports_i420(w: int, h: int, t: int) -> np.ndarray:
    """Fast motion and a scene cut (what a zero-centred +-12 search cannot follow): the background
    pans by (+37, -11) samples per picture, a 96x96 foreground object crosses it by (-45, +19) per
    picture, and at t = 3 the scene changes to other content altogether."""
    scene = 0 if t < 3 else 1
    ww, wh = w + 37 * 8, h + 11 * 8
    world = _world(300 + 10 * scene, ww, wh)
    ox, oy = (37 * t) % (ww - w + 1), (11 * (7 - t % 8)) % (wh - h + 1)
    Y = world[oy:oy + h, ox:ox + w].copy()
    obj = _world(900 + scene, 96, 96)
    px, py = (w - 96 - 45 * t) % max(w - 96, 1), (19 * t) % max(h - 96, 1)
    hh, ow = min(96, h - py), min(96, w - px)
    Y[py:py + hh, px:px + ow] = 255 - obj[:hh, :ow] // 2
    out = np.empty(i420_size(w, h), np.uint8)
    out[: w * h] = np.clip(Y, 0, 255).astype(np.uint8).ravel()
    c = Y[0::2, 0::2]
    out[w * h: w * h + w * h // 4] = (128 + (c % 32) - 16).astype(np.uint8).ravel()
    out[w * h + w * h // 4:] = (128 - (c % 24) + 12).astype(np.uint8).ravel()
    return out


def psnr(a: np.ndarray, b: np.ndarray) -> float:
    d = a.astype(np.float64) - b.astype(np.float64)
    mse = float(np.mean(d * d))
    return 99.0 if mse == 0 else 10.0 * np.log10(255.0 * 255.0 / mse)
