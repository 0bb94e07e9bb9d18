"""Pins oracle/input_pipeline.py against OpenCV (cv2, present in this image) and writes small golden vectors for the
GPU parity tests:  python oracle/make_input_golden.py  ->  tests/golden/input_pipeline.npz

The reference's transforms (virtex/factories.py:131-155) are albumentations objects that call cv2; albumentations is not
installed (no network), so the composition is rebuilt here from the SAME cv2 calls albumentations 1.x makes:
crop + cv2.resize(INTER_LINEAR), cv2.flip, cv2.LUT, cv2.cvtColor, cv2.addWeighted.
"""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import input_pipeline as P  # noqa: E402


def cv2_color_jitter(img, b, c, s, h, order):
    def brightness(x):
        return x if b == 1 else cv2.LUT(x, np.clip(np.arange(256, dtype=np.float64) * b, 0, 255).astype(np.uint8))

    def contrast(x):
        if c == 1:
            return x
        mean = cv2.cvtColor(x, cv2.COLOR_RGB2GRAY).mean()
        if c == 0:
            return np.full_like(x, int(mean + 0.5))
        return cv2.LUT(x, np.clip(np.arange(256, dtype=np.float64) * c + mean * (1 - c), 0, 255).astype(np.uint8))

    def saturation(x):
        if s == 1:
            return x
        gray = cv2.cvtColor(cv2.cvtColor(x, cv2.COLOR_RGB2GRAY), cv2.COLOR_GRAY2RGB)
        return gray if s == 0 else cv2.addWeighted(x, s, gray, 1 - s, 0)

    def hue(x):
        if h == 0:
            return x
        hsv = cv2.cvtColor(x, cv2.COLOR_RGB2HSV)
        lut = np.mod(np.arange(256, dtype=np.int16) + 180 * h, 180).astype(np.uint8)
        hsv[..., 0] = cv2.LUT(hsv[..., 0], lut)
        return cv2.cvtColor(hsv, cv2.COLOR_HSV2RGB)

    ops = (brightness, contrast, saturation, hue)
    for i in order:
        img = ops[i](np.ascontiguousarray(img))
    return img


def cv2_train(img, box, flip, jitter):
    y0, x0, h, w = box
    out = cv2.resize(np.ascontiguousarray(img[y0:y0 + h, x0:x0 + w]), (224, 224), interpolation=cv2.INTER_LINEAR)
    if flip:
        out = cv2.flip(out, 1)
    if jitter is not None:
        out = cv2_color_jitter(out, *jitter)
    return out


def cv2_val(img):
    nh, nw, oy, ox = P.val_geometry(*img.shape[:2])
    return cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR)[oy:oy + 224, ox:ox + 224]


def smooth_image(rng, H, W):
    """Natural-ish content (low-frequency colour field + noise) so that hue / saturation are not degenerate."""
    base = rng.uniform(0, 255, (H // 16 + 2, W // 16 + 2, 3)).astype(np.float32)
    img = cv2.resize(base, (W, H), interpolation=cv2.INTER_CUBIC) + rng.normal(0, 12, (H, W, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


def main():
    rng = np.random.default_rng(2024)
    sizes = [(480, 640), (427, 640), (333, 500), (120, 90), (224, 224), (640, 361)]
    images, boxes, flips, jitters, train_u8, val_u8 = [], [], [], [], [], []
    worst_hue = 0
    for k, (H, W) in enumerate(sizes):
        img = smooth_image(rng, H, W) if k % 2 == 0 else rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        box = P.sample_random_resized_crop(rng, H, W)
        flip = bool(rng.integers(0, 2))
        jit = P.sample_color_jitter(rng, p=0.85 if k else 1.0)
        ref = cv2_train(img, box, flip, jit)
        y0, x0, h, w = box
        mine = P.resize_linear_u8(img[y0:y0 + h, x0:x0 + w], 224, 224)
        assert np.array_equal(mine, cv2.resize(np.ascontiguousarray(img[y0:y0 + h, x0:x0 + w]), (224, 224),
                                               interpolation=cv2.INTER_LINEAR)), "resize not bit-exact"
        if flip:
            mine = mine[:, ::-1]
        if jit is not None:
            mine = P.color_jitter(mine, *jit)
        d = np.abs(mine.astype(int) - ref.astype(int))
        worst_hue = max(worst_hue, int(d.max()))
        assert d.max() <= 4 and (d > 0).mean() < 1e-3, (k, d.max(), (d > 0).mean())  # HSV2RGB float order, amplified by later ops
        assert np.array_equal(P.resize_linear_u8(img, *P.val_geometry(H, W)[:2])[
            P.val_geometry(H, W)[2]:P.val_geometry(H, W)[2] + 224, P.val_geometry(H, W)[3]:P.val_geometry(H, W)[3] + 224],
            cv2_val(img)) or min(H, W) < 224
        images.append(img)
        boxes.append(box)
        flips.append(flip)
        jitters.append(jit)
        train_u8.append(ref)
        if min(H, W) >= 224:
            val_u8.append(cv2_val(img))
    # primitives, exhaustively
    allrgb = rng.integers(0, 256, (256, 256, 3), dtype=np.uint8)
    assert np.array_equal(P.rgb2gray_u8(allrgb), cv2.cvtColor(allrgb, cv2.COLOR_RGB2GRAY))
    assert np.array_equal(P.rgb2hsv_u8(allrgb), cv2.cvtColor(allrgb, cv2.COLOR_RGB2HSV))
    hsv = cv2.cvtColor(allrgb, cv2.COLOR_RGB2HSV)
    dh = np.abs(P.hsv2rgb_u8(hsv).astype(int) - cv2.cvtColor(hsv, cv2.COLOR_HSV2RGB).astype(int))
    assert dh.max() <= 1 and (dh > 0).mean() < 1e-3, (dh.max(), (dh > 0).mean())
    for f in (0.6, 0.7, 1.3, 1.4):
        g = np.repeat(cv2.cvtColor(allrgb, cv2.COLOR_RGB2GRAY)[..., None], 3, -1)
        assert np.array_equal(P.add_weighted_u8(allrgb, f, g, 1 - f), cv2.addWeighted(allrgb, f, g, 1 - f, 0))
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "input_pipeline.npz")
    small = [0, 3, 4]  # keep the fixture small: three images with their cv2 outputs
    np.savez_compressed(
        out, n=len(small),
        **{f"img{i}": images[k] for i, k in enumerate(small)},
        **{f"box{i}": np.array(boxes[k]) for i, k in enumerate(small)},
        **{f"flip{i}": np.array(flips[k]) for i, k in enumerate(small)},
        **{f"jit{i}": np.array([-1.0] * 8 if jitters[k] is None else list(jitters[k][:4]) + list(jitters[k][4]))
           for i, k in enumerate(small)},
        **{f"train{i}": train_u8[k] for i, k in enumerate(small)})
    print("pinned against cv2", cv2.__version__, "| worst |oracle - cv2| over the jitter compositions:", worst_hue, "LSB ->", out,
          os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main()
