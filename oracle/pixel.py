"""Primitive-op CPU restatement of the pixel half of the hot path.  TEST INFRASTRUCTURE ONLY.

Library-free (numpy integer / float ops only, no cv2) restatement of

  * ``build_multich``                   Detect_OBB.py:87-133
  * ``dt_edge_channel_from_bgr``        Train_OBB.py:615-653
  * ``build_4ch_CHW_from_bgr_dtedge``   Train_OBB.py:655-664

at the bit level of what OpenCV 4.13.0 (IPP off) + numpy 2.3.5 compute for those calls
(SURVEY.md Appendix A).  Pinned: ``tests/test_oracle_pixel.py`` checks it against the
reference's own lifted functions on every tile of Input/Test1.png and Test2.png when
/root/reference is present, and against the committed vectors in ``tests/golden/``
everywhere else.

Every stage also returns its intermediate so the CUDA kernels can be compared stage by
stage (integer stages bit-exact).
"""
from __future__ import annotations

import math
from typing import Dict, Sequence

import numpy as np

HV = 62587          # round(0.955  * 2**16)  cv2 DIST_L2 3x3 axial weight
DG = 89738          # round(1.3693 * 2**16)  diagonal weight
DIST_MAX = 2**32 - 1 - DG      # cv2 4.13 saturation value; also what the out-of-image ring acts as
DEFAULT_SIGMAS = (0, 0.6, 1.2, 2.4)    # Detect_OBB.py:29


# ----------------------------------------------------------------------------- helpers

def reflect101(idx: np.ndarray, n: int) -> np.ndarray:
    """cv2 BORDER_REFLECT_101 index map (repeated reflection for |idx| >= n, 0 if n == 1)."""
    idx = np.asarray(idx, dtype=np.int64)
    if n == 1:
        return np.zeros_like(idx)
    period = 2 * (n - 1)
    m = np.mod(idx, period)
    return np.where(m >= n, period - m, m)


def gaussian_ksize(sigma: float) -> int:
    """``ksize = round(6*sigma + 1) | 1`` for 8-bit sources (cv2 GaussianBlur, ksize=(0,0))."""
    return int(round(sigma * 6 + 1)) | 1


def gaussian_kernel_q8(sigma: float) -> list:
    """8.8 fixed-point Gaussian taps as cv2 builds them for CV_8U (sum == 256).

    Float taps ``exp(-x^2/(2 sigma^2)) / sum`` are turned into integers by error diffusion
    from the outside in (round to nearest even of ``tap*256 + carried error``); the centre
    tap takes whatever is left of 256.
    """
    n = gaussian_ksize(sigma)
    half = n // 2
    scale = -0.5 * 0.25 / (sigma * sigma)
    vals = []
    x = 1 - n
    for _ in range(half):
        vals.append(math.exp((x * x) * scale))
        x += 2
    total = 2.0 * sum(vals) + 1.0
    inv = 1.0 / total
    taps = [0] * n
    err = 0.0
    acc = 0
    for i in range(half):
        adj = vals[i] * inv * 256.0 + err
        v = int(np.rint(adj))          # round half to even, like cvRound
        err = adj - v
        taps[i] = taps[n - 1 - i] = v
        acc += v
    taps[half] = 256 - 2 * acc
    return taps


def gray_u8(bgr: np.ndarray) -> np.ndarray:
    """cv2 BGR2GRAY for uint8: 15-bit fixed point (A.1)."""
    b = bgr[..., 0].astype(np.int64)
    g = bgr[..., 1].astype(np.int64)
    r = bgr[..., 2].astype(np.int64)
    return ((3735 * b + 19235 * g + 9798 * r + 16384) >> 15).astype(np.uint8)


def blur_u8(gray: np.ndarray, sigma: float) -> np.ndarray:
    """cv2.GaussianBlur(gray, (0,0), sigma, borderType=REFLECT_101) on uint8 (A.2)."""
    if sigma <= 0:
        return gray
    taps = gaussian_kernel_q8(sigma)
    r = len(taps) // 2
    h, w = gray.shape
    g = gray.astype(np.int64)
    cols = reflect101(np.arange(-r, w + r), w)
    gx = g[:, cols]
    hrow = np.zeros((h, w), dtype=np.int64)
    for i, k in enumerate(taps):
        hrow += k * gx[:, i:i + w]
    rows = reflect101(np.arange(-r, h + r), h)
    hy = hrow[rows, :]
    v = np.zeros((h, w), dtype=np.int64)
    for j, k in enumerate(taps):
        v += k * hy[j:j + h, :]
    return ((v + 32768) >> 16).astype(np.uint8)


def scharr_sq(img: np.ndarray) -> np.ndarray:
    """gx^2 + gy^2 of the 3x3 Scharr pair with REFLECT_101 borders, as exact integers (A.3)."""
    h, w = img.shape
    p = img.astype(np.int64)
    rows = reflect101(np.arange(-1, h + 1), h)
    cols = reflect101(np.arange(-1, w + 1), w)
    p = p[rows][:, cols]
    tl, tc, tr = p[:-2, :-2], p[:-2, 1:-1], p[:-2, 2:]
    ml, mr = p[1:-1, :-2], p[1:-1, 2:]
    bl, bc, br = p[2:, :-2], p[2:, 1:-1], p[2:, 2:]
    gx = 3 * (tr - tl) + 10 * (mr - ml) + 3 * (br - bl)
    gy = 3 * (bl - tl) + 10 * (bc - tc) + 3 * (br - tr)
    return gx * gx + gy * gy


def max_scharr_sq(gray: np.ndarray, sigmas: Sequence[float] = DEFAULT_SIGMAS) -> np.ndarray:
    """max over scales of gx^2+gy^2 (int64).  acc = sqrt_f32(f32(S)) is monotone in S."""
    S = None
    for s in sigmas:
        cur = scharr_sq(blur_u8(gray, s))
        S = cur if S is None else np.maximum(S, cur)
    return S


def acc_from_S(S: np.ndarray) -> np.ndarray:
    """cv2.magnitude (IPP off): one rounding of the integer sum to fp32, correctly rounded sqrt."""
    return np.sqrt(S.astype(np.float32))


def percentile_linear(values: np.ndarray, q: float) -> np.float64:
    """numpy ``percentile(method='linear')`` for one q on a flat array, float64 result (A.5)."""
    a = np.sort(values.reshape(-1), kind="stable")
    n = a.size
    vi = (n - 1) * (q / 100.0)
    lo = int(math.floor(vi))
    g = vi - lo
    a_lo, a_hi = a[lo], a[min(lo + 1, n - 1)]
    A = np.float64(a_lo)
    B = np.float64(a_hi)
    d = np.float64(a_hi - a_lo)        # numpy subtracts in the array dtype (fp32 here) before the lerp
    r = A + d * g
    if g >= 0.5:
        r = B - d * (1.0 - g)
    return np.float64(r)


def cross_open(mask: np.ndarray, iterations: int = 1) -> np.ndarray:
    """cv2.morphologyEx(MORPH_OPEN, 3x3 cross, iterations=n) (Detect_OBB.py:116-118, A.7): n erosions THEN n dilations
    (not n openings - an opening is idempotent), out-of-image neighbours ignored (erosion pads with "set", dilation
    with "clear").  Pinned on cv2 4.13 for n = 1, 2, 3 in tests/test_oracle_pixel.py."""
    m = mask.astype(bool)
    for _ in range(int(iterations)):
        e = np.pad(m, 1, constant_values=True)
        m = e[1:-1, 1:-1] & e[:-2, 1:-1] & e[2:, 1:-1] & e[1:-1, :-2] & e[1:-1, 2:]
    for _ in range(int(iterations)):
        d = np.pad(m, 1, constant_values=False)
        m = d[1:-1, 1:-1] | d[:-2, 1:-1] | d[2:, 1:-1] | d[1:-1, :-2] | d[1:-1, 2:]
    return m


def chamfer_fixed(zero_mask: np.ndarray) -> np.ndarray:
    """cv2.distanceTransform(DIST_L2, 3): 16.16 fixed-point 3x3 chamfer, two raster passes (A.8).

    ``zero_mask`` marks the zero pixels of the cv2 input (the opened edge pixels).  Returns
    the uint32 fixed-point field ``t``; the cv2 result is ``float32(t) * 2**-16``.
    Row-parallel form: 3-tap min from the previous row, then a min-plus prefix scan with
    slope HV along the row.  Every value saturates at DIST_MAX = UINT_MAX - DG, and the
    out-of-image ring never wins against an in-image zero pixel, so a tile with at least one
    zero pixel gets the plain chamfer distance and a tile with none gets DIST_MAX everywhere
    (65534.63 after scaling) - checked against cv2 4.13.0 with IPP off.
    """
    z = zero_mask.astype(bool)
    h, w = z.shape
    col = np.arange(w, dtype=np.int64) * HV
    f = np.empty((h, w), dtype=np.int64)
    up = np.full(w + 2, DIST_MAX, dtype=np.int64)
    for i in range(h):
        c = np.minimum(np.minimum(up[:-2] + DG, up[1:-1] + HV), up[2:] + DG)
        c[z[i]] = 0
        # the ring pixel at column -1 enters the scan with one HV step
        run = np.minimum.accumulate(np.minimum(c - col, DIST_MAX + HV)) + col
        run = np.minimum(run, DIST_MAX)
        f[i] = run
        up[1:-1] = run
    dn = np.full(w + 2, DIST_MAX, dtype=np.int64)
    for i in range(h - 1, -1, -1):
        c = np.minimum(np.minimum(dn[:-2] + DG, dn[1:-1] + HV), dn[2:] + DG)
        c = np.minimum(c, f[i])
        cr = c[::-1]
        run = np.minimum.accumulate(np.minimum(cr - col, DIST_MAX + HV)) + col
        run = np.minimum(run[::-1], DIST_MAX)
        f[i] = run
        dn[1:-1] = run
    return f.astype(np.uint32)


def chamfer_to_float(t: np.ndarray) -> np.ndarray:
    """``(float)(t * (1.f/65536))``: uint32 -> fp32 (round to nearest even) then exact scale."""
    return t.astype(np.float32) * np.float32(1.0 / 65536.0)


def normalize_minmax(acc: np.ndarray, dmin: float, dmax: float) -> np.ndarray:
    """cv2.normalize(acc, None, dmin, dmax, NORM_MINMAX) on fp32: scale / shift in float64, then
    convertTo's fma(src, f32(scale), f32(shift)) (A.9)."""
    smin = float(acc.min())
    smax = float(acc.max())
    scale = (dmax - dmin) * (1.0 / (smax - smin) if (smax - smin) > np.finfo(np.float64).eps else 0.0)
    shift = dmin - smin * scale
    fs = np.float32(scale)
    fh = np.float32(shift)
    # fp32 fma == round_f32(exact(src*fs + fh)); the 48-bit product is exact in float64
    prod = acc.astype(np.float64) * np.float64(fs)
    return _round_sum_to_f32(prod, np.float64(fh))


def normalize_minmax01(acc: np.ndarray) -> np.ndarray:
    """cv2.normalize(acc, None, 0, 1, NORM_MINMAX) (Detect_OBB.py:127)."""
    return normalize_minmax(acc, 0.0, 1.0)


def otsu_threshold_u8(img8: np.ndarray) -> int:
    """cv2.threshold(..., THRESH_OTSU) on 8-bit data: OpenCV's getThreshVal_Otsu_8u in float64
    (thresh.cpp; not in the reference tree - restated from the published algorithm and checked
    against cv2 4.13 in tests/test_oracle_pixel.py).  Returns the threshold; edges = img8 > thr."""
    h = np.bincount(img8.reshape(-1), minlength=256)
    scale = 1.0 / img8.size
    mu = 0.0
    for i in range(256):
        mu += float(i) * float(h[i])
    mu *= scale
    mu1 = q1 = max_sigma = 0.0
    max_val = 0
    eps = float(np.finfo(np.float32).eps)
    for i in range(256):
        p_i = float(h[i]) * scale
        mu1 *= q1
        q1 += p_i
        q2 = 1.0 - q1
        if min(q1, q2) < eps or max(q1, q2) > 1.0 - eps:
            continue
        mu1 = (mu1 + i * p_i) / q1
        mu2 = (mu - q1 * mu1) / q2
        sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2)
        if sigma > max_sigma:
            max_sigma = sigma
            max_val = i
    return max_val


def acc8_from_acc(acc: np.ndarray) -> np.ndarray:
    """``cv2.normalize(acc, None, 0, 255, NORM_MINMAX).astype(np.uint8)`` (Detect_OBB.py:110):
    fp32 result, then numpy's truncating cast."""
    return np.trunc(normalize_minmax(acc, 0.0, 255.0)).astype(np.int64).clip(0, 255).astype(np.uint8)


def _round_sum_to_f32(prod: np.ndarray, add: np.float64) -> np.ndarray:
    """round_to_f32(prod + add) for exact float64 ``prod``, immune to double rounding.

    ``s = prod + add`` is rounded to 53 bits first; that can only change the fp32 result
    when ``s`` lands exactly on an fp32 midpoint while the exact sum does not - the TwoSum
    error term then says on which side of the midpoint the exact sum lies.
    """
    s = prod + add
    bb = s - prod
    err = (prod - (s - bb)) + (add - bb)        # exact: true sum == s + err
    r = s.astype(np.float32)
    r64 = r.astype(np.float64)
    diff = s - r64
    toward = np.where(diff > 0, np.float32(np.inf), np.float32(-np.inf)).astype(np.float32)
    nb = np.nextafter(r, toward)
    is_tie = (diff != 0) & (np.abs(diff) * 2 == np.abs(nb.astype(np.float64) - r64))
    fix = is_tie & (np.sign(err) == np.sign(diff))
    return np.where(fix, nb, r).astype(np.float32)


# ----------------------------------------------------------------------------- full channel

def dt_edge_stages(bgr: np.ndarray, sigmas: Sequence[float] = DEFAULT_SIGMAS,
                   p_hi: float = 90.0, morph_open: int = 1, bin_method: str = "percentile") -> Dict[str, np.ndarray]:
    """All stages of the DT-Edge channel for one tile; ``bin_method`` = DT_BIN_METHOD
    ("percentile", the reference default, Detect_OBB.py:113-114; or "otsu", :109-111)."""
    gray = gray_u8(bgr)
    S = max_scharr_sq(gray, sigmas)
    acc = acc_from_S(S)
    if bin_method == "otsu":
        acc8 = acc8_from_acc(acc)
        hi = np.float64(otsu_threshold_u8(acc8))
        edges = acc8 > int(hi)
    else:
        hi = percentile_linear(acc, p_hi)
        edges = acc.astype(np.float64) >= hi
    opened = cross_open(edges, int(morph_open)) if int(morph_open) > 0 else edges
    t = chamfer_fixed(opened)
    dist = chamfer_to_float(t)
    lo1 = percentile_linear(dist, 1.0)
    hi99 = percentile_linear(dist, 99.0)
    d = np.clip((dist.astype(np.float64) - lo1) / max(1e-6, float(hi99 - lo1)), 0, 1)
    soft = np.exp(-d / 3.0)
    nrm = normalize_minmax01(acc)
    blend = 0.7 * soft + (np.float32(0.3) * nrm)
    blend = np.clip(blend, 0, 1)
    out = (blend * 255).astype(np.uint8)
    return {"gray": gray, "S": S, "acc": acc, "hi": hi, "edges": edges, "opened": opened,
            "t": t, "dist": dist, "p1": lo1, "p99": hi99, "nrm": nrm, "dt_edge": out}


def dt_edge_channel(bgr: np.ndarray, sigmas: Sequence[float] = DEFAULT_SIGMAS,
                    p_hi: float = 90.0, morph_open: int = 1, bin_method: str = "percentile") -> np.ndarray:
    return dt_edge_stages(bgr, sigmas, p_hi, morph_open, bin_method)["dt_edge"]


def build_multich(bgr: np.ndarray, out_channels: int = 3,
                  sigmas: Sequence[float] = DEFAULT_SIGMAS, bin_method: str = "percentile") -> np.ndarray:
    """3-ch: contiguous BGR copy; 4-ch: HWC [R,G,B,DT-Edge] uint8 (Detect_OBB.py:87-133)."""
    assert out_channels in (3, 4), f"Unsupported out_channels={out_channels}"
    if out_channels == 3:
        return np.ascontiguousarray(bgr)
    dt = dt_edge_channel(bgr, sigmas, bin_method=bin_method)
    return np.ascontiguousarray(np.dstack([bgr[..., 2], bgr[..., 1], bgr[..., 0], dt]).astype(np.uint8))


def build_4ch_chw(bgr: np.ndarray, sigmas: Sequence[float] = DEFAULT_SIGMAS) -> np.ndarray:
    """(4,H,W) variant (Train_OBB.py:655-664)."""
    return np.ascontiguousarray(build_multich(bgr, 4, sigmas).transpose(2, 0, 1))
