"""GPU input pipeline: the reference's per-sample CPU transforms + collate as four kernels per batch.

Drop-in for what `CaptioningDataset.__getitem__` + `collate_fn` (virtex/data/datasets/captioning.py:51-100) produce with
the transform lists of the base config (virtex/factories.py:131-155, `DATA.IMAGE_TRANSFORM_TRAIN/VAL`): the same batch
dict {"image" fp32 [B,3,224,224], "caption_tokens", "noitpac_tokens", "caption_lengths"}, built on the device from
decoded uint8 HWC images (any sizes) and token-id lists.  JPEG decoding, tokenisation and the caption-side
left<->right swap of the paired horizontal flip (virtex/data/transforms.py:29-36, a string operation) stay on the host;
the host also draws the random parameters, so the kernels are deterministic and testable:

    pipe = GpuInputPipeline(device)
    params = [pipe.sample_train_params(rng, *img.shape[:2]) for img in images]     # or pipe.val_params(H, W)
    batch = pipe(images, params, token_lists)

One pinned staging buffer and one H2D copy per batch carry the raw pixels (uint8: 4x fewer PCIe bytes than the fp32
tensors the reference's DataLoader ships) plus a ~100 B/image parameter table.  No CPU fallback: CUDA tensors out.
"""
import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .ops import _stream, call

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


class ImageParams:
    """Sampled parameters of one image: source region, resized size, output-window offset, flip, colour jitter."""
    __slots__ = ("region", "resized", "offset", "flip", "jitter")

    def __init__(self, region, resized, offset=(0, 0), flip=False, jitter=None):
        self.region, self.resized, self.offset, self.flip, self.jitter = region, resized, offset, flip, jitter


class GpuInputPipeline:
    def __init__(self, device, crop_size: int = 224, max_caption_length: int = 30, padding_idx: int = 0,
                 mean=IMAGENET_MEAN, std=IMAGENET_STD):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("GpuInputPipeline runs on a CUDA device (there is no CPU path)")
        self.S, self.max_len, self.pad = crop_size, max_caption_length, padding_idx
        m = np.array(mean, np.float32) * np.float32(255.0)
        inv = np.float32(1) / (np.array(std, np.float32) * np.float32(255.0))
        self.norm = torch.from_numpy(np.concatenate([m, inv])).to(self.device)
        self._pinned: Optional[torch.Tensor] = None
        self._dev: Optional[torch.Tensor] = None
        self._copied: Optional[torch.cuda.Event] = None  # the last H2D copy out of the pinned staging buffer

    # ------------------------------------------------------------------------------------------- host-side sampling
    def sample_train_params(self, rng: np.random.Generator, H: int, W: int, scale=(0.2, 1.0), ratio=(0.75, 1.333),
                            flip_p=0.5, jitter=(0.4, 0.4, 0.4, 0.1), jitter_p=0.8) -> ImageParams:
        """random_resized_crop -> horizontal_flip -> color_jitter of the base config (factories.py:136-152)."""
        area = H * W
        box = None
        for _ in range(10):
            target = rng.uniform(*scale) * area
            aspect = math.exp(rng.uniform(math.log(ratio[0]), math.log(ratio[1])))
            w, h = int(round(math.sqrt(target * aspect))), int(round(math.sqrt(target / aspect)))
            if 0 < w <= W and 0 < h <= H:
                box = (int(rng.integers(0, H - h + 1)), int(rng.integers(0, W - w + 1)), h, w)
                break
        if box is None:
            in_ratio = W / H
            if in_ratio < ratio[0]:
                w, h = W, int(round(W / ratio[0]))
            elif in_ratio > ratio[1]:
                h, w = H, int(round(H * ratio[1]))
            else:
                w, h = W, H
            box = ((H - h) // 2, (W - w) // 2, h, w)
        jit = None
        if rng.uniform() < jitter_p:
            b, c, s, hh = jitter
            jit = (rng.uniform(max(0, 1 - b), 1 + b), rng.uniform(max(0, 1 - c), 1 + c),
                   rng.uniform(max(0, 1 - s), 1 + s), rng.uniform(-hh, hh), tuple(int(i) for i in rng.permutation(4)))
        return ImageParams(box, (self.S, self.S), (0, 0), bool(rng.uniform() < flip_p), jit)

    def val_params(self, H: int, W: int, resize: int = 256) -> ImageParams:
        """smallest_resize(256) -> center_crop(224) (DATA.IMAGE_TRANSFORM_VAL)."""
        scale = resize / min(H, W)
        nh, nw = int(round(H * scale)), int(round(W * scale))
        if nh < self.S or nw < self.S:
            raise ValueError("image too small for the centre crop")
        return ImageParams((0, 0, H, W), (nh, nw), ((nh - self.S) // 2, (nw - self.S) // 2))

    # ------------------------------------------------------------------------------------------------------- batch
    def __call__(self, images: Sequence, params: Sequence[ImageParams],
                 token_lists: Optional[Sequence[Sequence[int]]] = None) -> Dict[str, torch.Tensor]:
        B, S = len(images), self.S
        assert B == len(params) and B > 0
        arrs = [np.ascontiguousarray(im.cpu().numpy() if torch.is_tensor(im) else im) for im in images]
        geom_i = np.zeros((B, 8), np.int32)
        geom_d = np.zeros((B, 2), np.float64)
        jit_i = np.zeros((B, 6), np.int32)
        jit_d = np.zeros((B, 4), np.float64)
        offs = np.zeros(B, np.int64)
        total = 0
        for n, (a, p) in enumerate(zip(arrs, params)):
            if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
                raise ValueError("images must be uint8 HWC RGB arrays")
            H, W = a.shape[:2]
            y0, x0, h, w = p.region
            if not (0 <= y0 and 0 <= x0 and y0 + h <= H and x0 + w <= W and h > 0 and w > 0):
                raise ValueError(f"crop box {p.region} outside the {H}x{W} image")
            geom_i[n] = (H, W, y0, x0, h, w, p.offset[0], p.offset[1])
            geom_d[n] = (h / p.resized[0], w / p.resized[1])
            jit_i[n, 0] = int(p.flip)
            jit_i[n, 2:] = (0, 1, 2, 3)
            jit_d[n] = (1.0, 1.0, 1.0, 0.0)
            if p.jitter is not None:
                jit_i[n, 1] = 1
                jit_d[n] = p.jitter[:4]
                jit_i[n, 2:] = p.jitter[4]
            offs[n] = total
            total += (a.size + 15) // 16 * 16
        tok_flat = tok_offs = None
        if token_lists is not None:
            assert len(token_lists) == B
            tok_offs = np.zeros(B + 1, np.int64)
            tok_offs[1:] = np.cumsum([len(t) for t in token_lists])
            tok_flat = np.fromiter((x for t in token_lists for x in t), np.int64, int(tok_offs[-1]))
        # ---- one pinned staging buffer: [pixels | tables], one H2D copy
        tables = [geom_d, jit_d, offs] + ([tok_flat, tok_offs] if tok_flat is not None else []) + [geom_i, jit_i]
        tab_bytes = sum((t.nbytes + 15) // 16 * 16 for t in tables)
        need = total + tab_bytes
        if self._pinned is None or self._pinned.numel() < need:
            self._pinned = torch.empty(int(need * 1.25) + 1024, dtype=torch.uint8).pin_memory()
            self._dev = torch.empty_like(self._pinned, device=self.device)
        if self._copied is not None:
            self._copied.synchronize()  # the previous batch's copy must have left the staging buffer
        host = self._pinned.numpy()
        for a, o in zip(arrs, offs):
            host[o:o + a.size] = a.reshape(-1)
        views, cur = [], total
        for t in tables:
            host[cur:cur + t.nbytes] = np.frombuffer(t.tobytes(), np.uint8)
            views.append((cur, t))
            cur += (t.nbytes + 15) // 16 * 16
        self._dev[:need].copy_(self._pinned[:need], non_blocking=True)
        self._copied = torch.cuda.Event()
        self._copied.record()
        base = self._dev.data_ptr()
        ptr = {id(t): base + o for o, t in views}
        s = _stream()
        img_u8 = torch.empty(B, S, S, 3, dtype=torch.uint8, device=self.device)
        gray = torch.zeros(B, dtype=torch.int64, device=self.device)
        out = torch.empty(B, 3, S, S, dtype=torch.float32, device=self.device)
        call("vtx_image_resample", base, ptr[id(offs)], ptr[id(geom_i)], ptr[id(geom_d)], ptr[id(jit_i)],
             img_u8.data_ptr(), B, S, s)
        call("vtx_image_gray_sum", img_u8.data_ptr(), ptr[id(jit_i)], ptr[id(jit_d)], gray.data_ptr(), B, S, s)
        call("vtx_image_jitter_normalize", img_u8.data_ptr(), ptr[id(jit_i)], ptr[id(jit_d)], gray.data_ptr(),
             self.norm.data_ptr(), out.data_ptr(), B, S, s)
        batch = {"image": out, "_image_u8": img_u8}
        if tok_flat is not None:
            T = int(min(self.max_len, max(len(t) for t in token_lists)))
            cap = torch.empty(B, T, dtype=torch.int64, device=self.device)
            rev = torch.empty(B, T, dtype=torch.int64, device=self.device)
            lens = torch.empty(B, dtype=torch.int64, device=self.device)
            call("vtx_collate_tokens", ptr[id(tok_flat)], ptr[id(tok_offs)], cap.data_ptr(), rev.data_ptr(),
                 lens.data_ptr(), B, T, self.max_len, self.pad, s)
            batch.update(caption_tokens=cap, noitpac_tokens=rev, caption_lengths=lens)
        return batch
