"""Seeded test images shared by the CPU (oracle pin) and GPU (parity) tests."""
import numpy as np


def blurred_noise(h, w, seed, passes=3):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, (h, w)).astype(np.float64)
    for _ in range(passes):
        a = (a + np.roll(a, 1, 0) + np.roll(a, -1, 0) + np.roll(a, 1, 1) + np.roll(a, -1, 1)) / 5.0
    a = (a - a.min()) / (a.max() - a.min() + 1e-9) * 255
    return a.astype(np.uint8)


def shapes(h, w, seed):
    """Rectangles, lines and noise patches on white: edges, grids and blobs."""
    rng = np.random.default_rng(seed)
    img = np.full((h, w), 255, np.uint8)
    for _ in range(12):
        x0, y0 = int(rng.integers(0, w - 4)), int(rng.integers(0, h - 4))
        x1, y1 = min(w, x0 + int(rng.integers(2, max(3, w // 3)))), min(h, y0 + int(rng.integers(2, max(3, h // 3))))
        img[y0:y1, x0:x1] = rng.integers(0, 256)
    for _ in range(6):
        y = int(rng.integers(0, h)); img[y, :] = rng.integers(0, 200)
        x = int(rng.integers(0, w)); img[:, x] = rng.integers(0, 200)
    ph, pw = max(2, h // 4), max(2, w // 4)
    y0, x0 = int(rng.integers(0, h - ph + 1)), int(rng.integers(0, w - pw + 1))
    img[y0:y0 + ph, x0:x0 + pw] = rng.integers(0, 256, (ph, pw))
    return img


def random_mask(h, w, seed, density=0.5):
    rng = np.random.default_rng(seed)
    return ((rng.random((h, w)) < density) * 255).astype(np.uint8)


def spiral_mask(h, w):
    """One long 1-px spiral: worst case for label propagation."""
    m = np.zeros((h, w), np.uint8)
    x0, y0, x1, y1 = 0, 0, w - 1, h - 1
    while x1 - x0 >= 2 and y1 - y0 >= 2:
        m[y0, x0:x1 + 1] = 255; m[y0:y1 + 1, x1] = 255
        m[y1, x0 + 2:x1 + 1] = 255; m[y0 + 2:y1 + 1, x0 + 2] = 255
        x0 += 4; y0 += 4; x1 -= 4; y1 -= 4
        if x1 - x0 >= 2 and y1 - y0 >= 2:
            m[y0 - 2, x0 - 2:x0 + 1] = 255
    return m


def checkerboard(h, w):
    yy, xx = np.mgrid[0:h, 0:w]
    return (((yy + xx) & 1) * 255).astype(np.uint8)


def rgb_noise(h, w, seed):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
