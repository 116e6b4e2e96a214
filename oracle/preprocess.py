"""TEST INFRASTRUCTURE (oracle): CPU restatement of the frame normalisation that precedes pnet:forward.

Reference call sites: load_image (utilities.lua:206-218: image.load + image.rgb2yuv for color_space = 'yuv'),
BatchIterator:processImage (BatchIterator.lua:146-161: per-channel centering, per-channel scaling by the unbiased
std when > 1e-8, then `img[1] = self.normalization:forward(img[{{1}}])` with
self.normalization = nn.SpatialContrastiveNormalization(1, image.gaussian1D(cfg.normalization.width)), :86).

`image` and `nn` are un-vendored Torch7 packages without a pinned version; their published algorithms
(2015) are restated here.  PARITY UNPINNED: there is no reference run or golden vector for this step.

  image.gaussian1D(size): sigma = 0.25, amplitude = 1, mean = 0.5, normalize = false:
      g[i] = exp(-(((i - center) / (sigma * size))^2) / 2), center = mean * size + 0.5, i = 1..size
  nn.SpatialSubtractiveNormalization(1, k1d): k = k1d / (sum(k1d) * sqrt(nInputPlane)); meanestimator = zero padding
      (size/2 each side) -> horizontal k -> vertical k; coef = meanestimator(ones); out = x - meanestimator(x) / coef
  nn.SpatialDivisiveNormalization(1, k1d, 1e-4, 1e-4): localstds = sqrt(meanestimator(x^2)); adjusted = localstds / coef;
      thresholded = adjusted > 1e-4 ? adjusted : 1e-4; out = x / thresholded
  nn.SpatialContrastiveNormalization = Subtractive then Divisive, same kernel."""
import numpy as np
import torch
import torch.nn.functional as TF


def gaussian1D(size, sigma=0.25, amplitude=1.0, mean=0.5):
    center = mean * size + 0.5
    i = np.arange(1, size + 1, dtype=np.float64)
    return (amplitude * np.exp(-(((i - center) / (sigma * size)) ** 2) / 2.0)).astype(np.float32)


def rgb2yuv(img):
    """image.rgb2yuv on a [3][H][W] float tensor."""
    r, g, b = img[0], img[1], img[2]
    y = 0.299 * r + 0.587 * g + 0.114 * b
    u = -0.14713 * r - 0.28886 * g + 0.436 * b
    v = 0.615 * r - 0.51499 * g - 0.10001 * b
    return torch.stack([y, u, v]).to(torch.float32)


def _mean_estimator(x, k):
    """zero padding -> horizontal k -> vertical k on a [H][W] plane."""
    r = len(k) // 2
    kt = torch.from_numpy(k)
    t = TF.conv2d(x[None, None], kt.view(1, 1, 1, -1), padding=(0, r))
    t = TF.conv2d(t, kt.view(1, 1, -1, 1), padding=(r, 0))
    return t[0, 0]


def contrastive_normalization(plane, width=7, threshold=1e-4):
    k = gaussian1D(width)
    k = (k / np.float32(k.sum())).astype(np.float32)
    coef = _mean_estimator(torch.ones_like(plane), k)
    s = plane - _mean_estimator(plane, k) / coef
    sd = torch.sqrt(_mean_estimator(s * s, k)) / coef
    sd = torch.where(sd > threshold, sd, torch.full_like(sd, threshold))
    return s / sd


def normalize_frame(img, rgb_to_yuv=False, centering=True, scaling=True, contrastive_width=7):
    """img: [3][H][W] float32 tensor -> normalised copy (BatchIterator.lua:146-161 order)."""
    x = img.clone().to(torch.float32)
    if rgb_to_yuv:
        x = rgb2yuv(x)
    if centering:
        for i in range(3):
            x[i] = x[i] - np.float32(x[i].double().mean().item())       # TH meanall accumulates in double
    if scaling:
        for i in range(3):
            s = x[i].double().std(unbiased=True).item()                  # TH stdall, double accumulation
            if s > 1e-8:
                x[i] = x[i] / np.float32(s)
    if contrastive_width:
        x[0] = contrastive_normalization(x[0], contrastive_width)
    return x


# ---------------------------------------------------------------------------------------------- resize
def find_target_size(orig_w, orig_h, target_smaller_side, max_pixel_size):
    """utilities.lua:188-204 (Lua numbers are doubles; math.floor(x + 0.5))."""
    import math
    if orig_h < orig_w:
        w = min(orig_w * target_smaller_side / orig_h, max_pixel_size)
        h = math.floor(orig_h * w / orig_w + 0.5)
        w = math.floor(w + 0.5)
    else:
        h = min(orig_h * target_smaller_side / orig_w, max_pixel_size)
        w = math.floor(orig_w * h / orig_h + 0.5)
        h = math.floor(h + 0.5)
    assert w >= 1 and h >= 1
    return int(w), int(h)


def _scale_linear_1d(src, dst_len):
    """image's Main_scaleLinear_rowcol along the LAST axis of a float32 array (torch/image generic/image.c, 2015; un-vendored,
    restated from the published source -- PARITY UNPINNED): enlarging = linear interpolation between the two neighbours with
    scale (src_len - 1) / (dst_len - 1), the last sample copied; shrinking = the mean of the source interval
    [di * scale, (di + 1) * scale) with fractional weights at both ends, scale = src_len / dst_len; float32 arithmetic in
    the C loop's order."""
    f32 = np.float32
    src = np.ascontiguousarray(src, dtype=f32)
    src_len = src.shape[-1]
    dst = np.empty(src.shape[:-1] + (dst_len,), dtype=f32)
    if dst_len > src_len:
        if src_len == 1:
            dst[...] = src
            return dst
        scale = f32(src_len - 1) / f32(dst_len - 1)
        for di in range(dst_len - 1):
            si_f = f32(di) * scale
            si_i = int(si_f)
            si_f = f32(si_f - f32(si_i))
            dst[..., di] = f32(1) * (f32(1) - si_f) * src[..., si_i] + si_f * src[..., si_i + 1]
        dst[..., dst_len - 1] = src[..., src_len - 1]
    elif dst_len < src_len:
        scale = f32(src_len) / f32(dst_len)
        si0_i, si0_f = 0, f32(0)
        for di in range(dst_len):
            si1_f = f32(di + 1) * scale
            si1_i = int(si1_f)
            si1_f = f32(si1_f - f32(si1_i))
            acc = (f32(1) - si0_f) * src[..., si0_i]
            n = f32(1) - si0_f
            for si in range(si0_i + 1, si1_i):
                acc = acc + src[..., si]
                n = f32(n + f32(1))
            if si1_i < src_len:
                acc = acc + si1_f * src[..., si1_i]
                n = f32(n + si1_f)
            dst[..., di] = acc / n
            si0_i, si0_f = si1_i, si1_f
    else:
        dst[...] = src
    return dst


def scale_image(img, width, height):
    """image.scale(img, width, height) in its default 'bilinear' mode on a [C][H][W] float32 array: rows first (width) into
    a [C][H][width] temporary, then columns (BatchIterator.lua:49-52 via transform_example)."""
    a = np.asarray(img, dtype=np.float32)
    tmp = _scale_linear_1d(a, int(width))
    out = _scale_linear_1d(np.swapaxes(tmp, -1, -2), int(height))
    return np.ascontiguousarray(np.swapaxes(out, -1, -2))
