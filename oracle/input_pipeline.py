"""CPU oracle (TEST INFRASTRUCTURE -- never imported by the product) of the reference's image/caption input pipeline,
the SURVEY section 8 row f-3 "next" component:

  virtex/data/datasets/captioning.py:51-100   __getitem__ (image_transform, HWC -> CHW, [SOS] .. [EOS], trim) + collate_fn
  virtex/factories.py:131-155                 random_resized_crop(scale=(0.2, 1), ratio=(0.75, 1.333)), horizontal_flip(0.5),
                                              color_jitter(0.4, 0.4, 0.4, 0.1, p=0.8), normalize(ImageNet mean / std),
                                              smallest_resize(256) + center_crop(224) for validation
  virtex/data/transforms.py:5-40              HorizontalFlip = cv2.flip(img, 1)

The arithmetic lives in two third-party dependencies of the reference: OpenCV (present in this image, cv2 4.13) and
albumentations >= 1.0 (requirements.txt; ABSENT here, no network).  What is restated, in numpy integer / float32 ops:

  * cv2.resize(..., INTER_LINEAR) on uint8 (what albumentations' RandomResizedCrop / SmallestMaxSize call): 11-bit
    fixed-point coefficients, horizontal pass to int32, vertical pass `((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2) >> 2`
    -- PINNED: bit-exact against cv2.resize itself (oracle/make_input_golden.py, tests/test_input_pipeline.py);
  * cv2.cvtColor RGB2GRAY (15-bit fixed point) and RGB2HSV (12-bit division tables) -- PINNED bit-exact against cv2;
    HSV2RGB (float32, truncating store) -- pinned to cv2 within 1 LSB on < 0.01 % of values (float evaluation order);
  * cv2.addWeighted on uint8 = round-half-even(fmaf(a, alpha, b * beta)) in float32 -- PINNED bit-exact against cv2;
  * albumentations 1.x `ColorJitter` (brightness / contrast / saturation / hue as uint8 LUTs + addWeighted + HSV hue
    shift, applied in a random order) and `Normalize` ((img/255 - mean)/std in float32): restated from the published
    algorithm (albumentations/augmentations/functional.py `adjust_*_torchvision`), built from the cv2-pinned primitives
    above.  albumentations itself cannot be imported here: parity of the COMPOSITION is "unpinned" and the tests say so.

Random parameter SAMPLING (crop box, flip coin, jitter factors, op order) is host logic outside the kernels: both the
oracle and the product take the sampled parameters as inputs.
"""
import math

import numpy as np

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


# ------------------------------------------------------------------------------------------------ cv2.resize (uint8)
def _coeffs(dn, sn, scale, offset, clamp_f):
    """Source index and 11-bit coefficient pair for every destination index (cv::resize linear, uint8).
    `scale` / `offset`: source = offset + (d + 0.5) * scale - 0.5 (offset != 0 expresses a crop)."""
    idx = np.empty(dn, np.int64)
    a = np.empty((dn, 2), np.int64)
    for d in range(dn):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(math.floor(f))
        f = np.float32(f - np.float32(s))
        if clamp_f:  # x direction: the fraction is zeroed at the borders; y direction: only the row index is clamped
            if s < 0:
                s, f = 0, np.float32(0)
            if s >= sn - 1:
                s, f = sn - 1, np.float32(0)
        idx[d] = s
        a[d, 0] = int(np.rint(np.float32((np.float32(1) - f) * np.float32(2048))))
        a[d, 1] = int(np.rint(np.float32(f * np.float32(2048))))
    return idx, a


def resize_linear_u8(src, dh, dw):
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR) for uint8 [H, W, C]."""
    sh, sw = src.shape[:2]
    xi, xa = _coeffs(dw, sw, sw / dw, 0, True)
    yi, ya = _coeffs(dh, sh, sh / dh, 0, False)
    S = src.astype(np.int64)
    x1 = np.minimum(xi + 1, sw - 1)
    Hp = S[:, xi] * xa[:, 0][None, :, None] + S[:, x1] * xa[:, 1][None, :, None]
    y0, y1 = np.clip(yi, 0, sh - 1), np.clip(yi + 1, 0, sh - 1)
    b0, b1 = ya[:, 0][:, None, None], ya[:, 1][:, None, None]
    out = (((b0 * (Hp[y0] >> 4)) >> 16) + ((b1 * (Hp[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


# ------------------------------------------------------------------------------------------------ cv2 colour ops
def rgb2gray_u8(img):
    r, g, b = (img[..., i].astype(np.int64) for i in range(3))
    return ((r * 9798 + g * 19235 + b * 3735 + (1 << 14)) >> 15).astype(np.uint8)


def _div_table(num, i, div):
    out = np.zeros_like(i)
    nz = i > 0
    out[nz] = np.rint((num << 12) / (div * i[nz].astype(np.float64))).astype(np.int64)
    return out


def rgb2hsv_u8(img):
    """cv2.cvtColor(img, COLOR_RGB2HSV) for uint8: H in [0, 180), S and V in [0, 255]."""
    r, g, b = (img[..., i].astype(np.int64) for i in range(3))
    v = np.maximum(np.maximum(r, g), b)
    diff = v - np.minimum(np.minimum(r, g), b)
    s = (diff * _div_table(255, v, 1.0) + (1 << 11)) >> 12
    h = np.where(v == r, g - b, np.where(v == g, b - r + 2 * diff, r - g + 4 * diff))
    h = (h * _div_table(180, diff, 6.0) + (1 << 11)) >> 12
    h = h + np.where(h < 0, 180, 0)
    return np.stack([h, s, v], -1).astype(np.uint8)


def hsv2rgb_u8(hsv):
    """cv2.cvtColor(hsv, COLOR_HSV2RGB) for uint8 (float32 evaluation, truncating store)."""
    one = np.float32(1)
    h = hsv[..., 0].astype(np.float32) * np.float32(6.0 / 180.0)
    s = hsv[..., 1].astype(np.float32) * np.float32(1 / 255.0)
    v = hsv[..., 2].astype(np.float32) * np.float32(1 / 255.0)
    sector = np.floor(h).astype(np.int64)
    f = h - sector.astype(np.float32)
    sector = sector % 6
    tab = np.stack([v, v * (one - s), v * (one - s * f), v * (one - s * (one - f))], -1)
    sd = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]])
    bgr = np.take_along_axis(tab, sd[sector], -1)
    return np.clip(np.floor(bgr[..., ::-1] * np.float32(255.0)), 0, 255).astype(np.uint8)


def add_weighted_u8(a, alpha, b, beta):
    """cv2.addWeighted(a, alpha, b, beta, 0) for uint8: float32 fma, round half to even, saturate."""
    al, be = np.float32(alpha), np.float32(beta)
    t = b.astype(np.float32) * be
    v = (a.astype(np.float64) * np.float64(al) + t.astype(np.float64)).astype(np.float32)  # == fmaf(a, al, t)
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


# ------------------------------------------------------------------------------------------------ albumentations ops
def _lut(scale, bias):
    """uint8 LUT `clip(arange(256) * scale + bias, 0, 255).astype(uint8)` (float64 like numpy's default; truncation)."""
    return np.clip(np.arange(256, dtype=np.float64) * scale + bias, 0, 255).astype(np.uint8)


def adjust_brightness(img, factor):
    if factor == 1:
        return img
    return _lut(factor, 0.0)[img]


def contrast_mean(img):
    """Mean of the grey image as albumentations takes it (cv2 mean of the uint8 grey image, a double)."""
    return float(rgb2gray_u8(img).astype(np.float64).mean())


def adjust_contrast(img, factor, mean=None):
    if factor == 1:
        return img
    mean = contrast_mean(img) if mean is None else mean
    if factor == 0:
        return np.full_like(img, int(mean + 0.5))
    return _lut(factor, mean * (1 - factor))[img]


def adjust_saturation(img, factor):
    if factor == 1:
        return img
    gray = np.repeat(rgb2gray_u8(img)[..., None], 3, -1)
    if factor == 0:
        return gray
    return add_weighted_u8(img, factor, gray, 1 - factor)


def adjust_hue(img, factor):
    if factor == 0:
        return img
    hsv = rgb2hsv_u8(img)
    lut = np.mod(np.arange(256, dtype=np.int16) + 180 * factor, 180).astype(np.uint8)
    hsv[..., 0] = lut[hsv[..., 0]]
    return hsv2rgb_u8(hsv)


def color_jitter(img, brightness, contrast, saturation, hue, order):
    """albumentations.ColorJitter.apply: the four ops in the sampled `order` (a permutation of 0..3)."""
    ops = (lambda x: adjust_brightness(x, brightness), lambda x: adjust_contrast(x, contrast),
           lambda x: adjust_saturation(x, saturation), lambda x: adjust_hue(x, hue))
    for i in order:
        img = ops[i](img)
    return img


def normalize_chw(img, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """albumentations.Normalize (max_pixel_value 255) followed by HWC -> CHW (captioning.py:66)."""
    m = np.array(mean, np.float32) * np.float32(255.0)
    inv = np.float32(1) / (np.array(std, np.float32) * np.float32(255.0))
    out = (img.astype(np.float32) - m) * inv
    return np.ascontiguousarray(out.transpose(2, 0, 1))


# ------------------------------------------------------------------------------------------------ whole transforms
def train_transform(img, box, flip, jitter, size=224):
    """IMAGE_TRANSFORM_TRAIN of the base config: random_resized_crop -> horizontal_flip -> color_jitter -> normalize.
    box = (y0, x0, h, w) sampled crop; flip: bool; jitter: None or (brightness, contrast, saturation, hue, order)."""
    y0, x0, h, w = box
    out = resize_linear_u8(img[y0:y0 + h, x0:x0 + w], size, size)
    if flip:
        out = out[:, ::-1]
    if jitter is not None:
        out = color_jitter(out, *jitter)
    return normalize_chw(out)


def val_geometry(H, W, resize=256, crop=224):
    """smallest_resize(256) + center_crop(224): resized size and crop offsets (albumentations rounding rules)."""
    scale = resize / min(H, W)
    nh, nw = int(round(H * scale)), int(round(W * scale))  # py3 round: half to even, as albumentations' py3_round
    return nh, nw, (nh - crop) // 2, (nw - crop) // 2


def val_transform(img, resize=256, crop=224):
    """IMAGE_TRANSFORM_VAL: smallest_resize -> center_crop -> normalize."""
    nh, nw, oy, ox = val_geometry(img.shape[0], img.shape[1], resize, crop)
    out = resize_linear_u8(img, nh, nw)[oy:oy + crop, ox:ox + crop]
    return normalize_chw(out)


def sample_random_resized_crop(rng, H, W, scale=(0.2, 1.0), ratio=(0.75, 1.333)):
    """albumentations.RandomResizedCrop.get_params_dependent_on_targets: (y0, x0, h, w)."""
    area = H * W
    for _ in range(10):
        target = rng.uniform(*scale) * area
        aspect = math.exp(rng.uniform(math.log(ratio[0]), math.log(ratio[1])))
        w, h = int(round(math.sqrt(target * aspect))), int(round(math.sqrt(target / aspect)))
        if 0 < w <= W and 0 < h <= H:
            return int(rng.integers(0, H - h + 1)), int(rng.integers(0, W - w + 1)), h, w
    in_ratio = W / H
    if in_ratio < ratio[0]:
        w, h = W, int(round(W / ratio[0]))
    elif in_ratio > ratio[1]:
        h, w = H, int(round(H * ratio[1]))
    else:
        w, h = W, H
    return (H - h) // 2, (W - w) // 2, h, w


def sample_color_jitter(rng, brightness=0.4, contrast=0.4, saturation=0.4, hue=0.1, p=0.8):
    if rng.uniform() >= p:
        return None
    return (rng.uniform(max(0, 1 - brightness), 1 + brightness), rng.uniform(max(0, 1 - contrast), 1 + contrast),
            rng.uniform(max(0, 1 - saturation), 1 + saturation), rng.uniform(-hue, hue), tuple(rng.permutation(4)))


def collate_captions(token_lists, max_len=30, pad=0):
    """captioning.py:68-100: trim to max_len, right-pad with `pad` to the longest caption of the batch, and the
    reversed copy ("noitpac") padded the same way."""
    toks = [list(t)[:max_len] for t in token_lists]
    T = max(len(t) for t in toks)
    cap = np.full((len(toks), T), pad, np.int64)
    rev = np.full((len(toks), T), pad, np.int64)
    for i, t in enumerate(toks):
        cap[i, :len(t)] = t
        rev[i, :len(t)] = t[::-1]
    return cap, rev, np.array([len(t) for t in toks], np.int64)
