"""Image generators with other statistics than gslam_b200.synth (ties, saturation, flat regions) for the ORB parity tests."""
import numpy as np


def image_of_kind(kind, w, h, seed):
    r = np.random.default_rng(seed)
    if kind == "noise":
        return r.integers(0, 256, (h, w), dtype=np.uint8)
    if kind == "binary":
        return (r.integers(0, 2, (h, w)) * 255).astype(np.uint8)
    if kind == "checker":  # exact ties everywhere: FAST scores, Harris responses, moments
        c = int(r.integers(3, 17)); yy, xx = np.mgrid[0:h, 0:w]
        return ((((yy // c) + (xx // c)) & 1) * 255).astype(np.uint8)
    if kind == "blurred":
        import cv2
        return cv2.GaussianBlur(r.integers(0, 256, (h, w), dtype=np.uint8), (0, 0), 1.7)
    if kind == "blocks":
        b = int(r.integers(4, 24)); small = r.integers(0, 256, ((h + b - 1) // b, (w + b - 1) // b), dtype=np.uint8)
        return np.kron(small, np.ones((b, b), np.uint8))[:h, :w].copy()
    return (128 + r.integers(-12, 13, (h, w))).astype(np.uint8)  # low contrast


KIND_CASES = [
    ("noise", 333, 257, 1, 3000, 8, 1.2, 20), ("binary", 201, 315, 2, 500, 5, 1.33, 40), ("checker", 412, 300, 3, 500, 8, 1.2, 20),
    ("blurred", 640, 199, 4, 150, 3, 1.5, 5), ("blocks", 275, 275, 5, 9, 8, 1.1, 20), ("lowcontrast", 500, 400, 6, 2, 1, 1.2, 5),
    ("checker", 96, 131, 7, 1, 8, 1.2, 5)]
