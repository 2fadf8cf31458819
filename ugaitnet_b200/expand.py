"""Host side of the device-side missing-modality expansion.

The reference's generator builds the E-fold expanded batch on the CPU and ships all of it to the GPU
(/root/reference/data/mj_dataGeneratorMMUWYHsingle.py:780-812).  Here only the PATTERN is produced on
the host -- a [B0*E, M] 0/1 table, drawn with Python's `random` exactly as the reference does -- and the
volumes are expanded by `ugn_pack_input_expand` while they are packed, so only the B0 base rows cross PCIe.
"""
from __future__ import annotations

import random as _random
from typing import Optional, Tuple

import numpy as np

NOISE = 1e-9      # data/mj_dataGeneratorMMUWYHsingle.py:102


def expansion_pattern(n_base: int, expand: int, nmods: int, rng: Optional[_random.Random] = None
                      ) -> Tuple[np.ndarray, np.ndarray]:
    """Returns (src_row i32 [B], use f32 [B, nmods]) with B = n_base*max(expand,1).

    Row i*E keeps every modality; rows i*E+1+ex follow :791-803 -- even i: min(ex+1, M-1) draws (with
    replacement, `random.randrange`) of a modality to disable (E == 2: the number of draws itself is
    drawn from [1, M)); odd i: only modality (i+ex) % 3 enabled."""
    rng = rng or _random
    E = max(int(expand), 1)
    src = np.repeat(np.arange(n_base, dtype=np.int32), E)
    use = np.ones((n_base * E, nmods), dtype=np.float32)
    for i in range(n_base):
        for ex in range(E - 1):
            if i % 2 == 0:
                ndis = min(ex + 1, nmods - 1) if E > 2 else rng.randrange(1, nmods, 1)
                l_dis = [1] * nmods
                for _ in range(ndis):
                    l_dis[rng.randrange(0, nmods, 1)] = 0
            else:
                l_dis = [0] * nmods
                l_dis[(i + ex) % 3 % nmods] = 1
            use[i * E + ex + 1] = l_dis
    return src, use


def expand_on_host(base, src, use_col, noise: float = NOISE):
    """numpy restatement of what the device pack does for one modality (tests / documentation)."""
    out = np.asarray(base)[src].copy()
    out[use_col == 0] = noise
    return out


def mirror_sequence(sample: np.ndarray) -> np.ndarray:
    """data/mj_augmentation.py:12-32 restated: sample [C,H,W]; every channel flipped left-right, even
    channels negated (for every modality -- the `isof` argument of the reference is never read)."""
    out = np.asarray(sample)[..., ::-1].copy()
    out[0::2] = -out[0::2]
    return out


def to_gaitset_layout(sample: np.ndarray) -> np.ndarray:
    """Stacked-frame sample [C,H,W] -> the gaitset=True input [25,H,W,c]
    (data/mj_dataGeneratorMMUWYHsingle_repetitions.py:426-434): an optical-flow sample (C == 50, channels interleaved
    x0,y0,x1,y1,...) becomes 25 frames of 2 channels (x = even, y = odd channels), any other sample [T,H,W] gets a
    trailing channel axis of 1.  Accepts a leading batch axis as well."""
    x = np.asarray(sample)
    if x.ndim == 4:
        return np.stack([to_gaitset_layout(s) for s in x])
    if x.shape[0] == 50:
        out = np.zeros((25, x.shape[1], x.shape[2], 2), dtype=x.dtype)
        out[..., 0] = x[::2]
        out[..., 1] = x[1::2]
        return out
    return x[..., None].copy()


def shift_sequence(sample: np.ndarray, tx: int, ty: int) -> np.ndarray:
    """mj_transformsequence (data/mj_augmentation.py:35-50) for a transform that is a pure integer displacement
    {tx, ty} (ImageDataGenerator width/height_shift_range = [-5,-3,0,3,5], :62-64): every frame goes through
    scipy.ndimage.affine_transform(frame, identity, offset=(tx, ty), order=1, mode='nearest'), i.e.
    out[y, x] = in[clamp(y + tx), clamp(x + ty)] (pinned against scipy in tests/test_dist_cpu.py).  sample [C,H,W]."""
    x = np.asarray(sample)
    H, W = x.shape[-2:]
    ys = np.clip(np.arange(H) + int(tx), 0, H - 1)
    xs = np.clip(np.arange(W) + int(ty), 0, W - 1)
    return x[..., ys[:, None], xs[None, :]].copy()


def clip_flow(sample: np.ndarray, lo: float = 0.05, hi: float = 2.3, val: float = 1e-11) -> np.ndarray:
    """The optical-flow magnitude clip of __load_dd (data/mj_dataGeneratorMMUWYHsingle.py:318-324) on DECODED values:
    raw |v| > 2300 or < 50 -> 1e-8, then / compressFactor (100) * 0.1."""
    x = np.asarray(sample).copy()
    a = np.abs(x)
    x[(a > hi) | (a < lo)] = val
    return x


def augment_on_host(base, src, use_col, mirror=None, shift=None, clip=None, noise: float = NOISE):
    """numpy restatement of ugn_pack_input_augment for one modality: clip -> integer shift -> mirror -> expansion."""
    base = np.asarray(base)
    out = np.empty((len(src),) + base.shape[1:], dtype=base.dtype)
    for i, s in enumerate(src):
        v = base[s]
        if clip is not None and clip[i]:
            v = clip_flow(v)
        if shift is not None:
            v = shift_sequence(v, int(shift[i][0]), int(shift[i][1]))
        if mirror is not None and mirror[i]:
            v = mirror_sequence(v)
        out[i] = v if use_col[i] != 0 else noise
    return out
