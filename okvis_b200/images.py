"""Seeded synthetic 752x480 grey images for the frontend configuration (BASELINE.json configs[2]):
a textured scene (random rectangles, discs and a low-frequency shading) and a second view of it with a
horizontal disparity and sensor noise."""
import numpy as np


def textured_image(seed=0x0B200 + 3000, width=752, height=480, n_shapes=900):
    rng = np.random.Generator(np.random.PCG64(seed))
    img = np.full((height, width), 110.0)
    yy, xx = np.mgrid[0:height, 0:width]
    img += 25.0 * np.sin(xx / 97.0) * np.cos(yy / 71.0)
    for _ in range(n_shapes):
        cx, cy = rng.uniform(0, width), rng.uniform(0, height)
        w, h = rng.uniform(4, 40), rng.uniform(4, 40)
        val = rng.uniform(20, 235)
        if rng.random() < 0.6:
            x0, x1 = int(max(cx - w / 2, 0)), int(min(cx + w / 2, width))
            y0, y1 = int(max(cy - h / 2, 0)), int(min(cy + h / 2, height))
            img[y0:y1, x0:x1] = val
        else:
            m = (xx - cx) ** 2 + (yy - cy) ** 2 < (w / 2) ** 2
            img[m] = val
    img += rng.normal(0, 1.5, img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def stereo_pair(seed=0x0B200 + 3000, disparity=12, noise_sigma=2.0):
    left = textured_image(seed)
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    right = np.roll(left, -disparity, axis=1).astype(np.float64)
    right[:, -disparity:] = left[:, -disparity:]
    right += rng.normal(0, noise_sigma, right.shape)
    return left, np.clip(np.rint(right), 0, 255).astype(np.uint8)
