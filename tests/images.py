"""Synthetic inputs. Generators follow the reference's own test/bench images (cited per function)."""
import numpy as np


def ref_img_rgb(width=258, height=128):
    """create_test_img_rgb, /root/reference/src/lib.rs:81-98 (258 wide => odd MCU count)."""
    y, x = np.mgrid[0:height, 0:width]
    x = np.minimum(x, 255)
    img = np.stack([x, y * 2, (x + y * 2) // 2], axis=-1)
    return (img & 0xFF).astype(np.uint8)


def ref_img_rgba(width=258, height=128):
    """create_test_img_rgba, src/lib.rs:100-118."""
    rgb = ref_img_rgb(width, height)
    x = np.minimum(np.mgrid[0:height, 0:width][1], 255).astype(np.uint8)
    return np.concatenate([rgb, x[..., None]], axis=-1)


def ref_img_gray(width=258, height=128):
    """create_test_img_gray, src/lib.rs:120-135 (luma of the rgb test image, reference formula)."""
    rgb = ref_img_rgb(width, height).astype(np.int32)
    yv = (19595 * rgb[..., 0] + 38470 * rgb[..., 1] + 7471 * rgb[..., 2] + 0x7FFF) >> 16
    return yv.astype(np.uint8)


def ref_img_cmyk(width=258, height=192):
    """create_test_img_cmyk, src/lib.rs:137-154."""
    y, x = np.mgrid[0:height, 0:width]
    x = np.minimum(x, 255)
    img = np.stack([x, y * 3 // 2, (x + y * 3 // 2) // 2, 255 - (x + y) // 2], axis=-1)
    return (img & 0xFF).astype(np.uint8)


def bench_img(width=2000, height=1800):
    """create_bench_img, /root/reference/criterion/benches/encode.rs:6-55, at any size."""
    y, x = np.mgrid[0:height, 0:width].astype(np.int64)
    xy = x * y
    img = np.stack([x % 256, x % 256, xy % 256], axis=-1).astype(np.uint8)
    rules = [(29, (96, 96, 255)), (27, (255, 96, 96)), (25, (96, 255, 96)), (23, (0, 255, 0)),
             (21, (0, 0, 255)), (19, (255, 0, 0)), (17, (255, 255, 255)), (13, (0, 0, 0))]
    for m, c in rules:  # later rules have priority (reversed if/else chain)
        img[xy % m == 0] = c
    return img


def lcg_bytes(n, seed=42):
    """SimpleRng, /root/reference/src/avx2/ycbcr.rs:164-190: state = state*6364136223846793005 + 1, low byte."""
    out = np.empty(n, dtype=np.uint8)
    state = seed
    mask = (1 << 64) - 1
    for i in range(n):
        state = (state * 6364136223846793005 + 1) & mask
        out[i] = state & 0xFF
    return out


def photo_like(width, height, channels=3, seed=42):
    """Smooth gradients + band-limited texture + a little noise: compresses like a photograph.
    Deterministic (numpy PCG64 with fixed seed), cheap at 16384^2."""
    rng = np.random.default_rng(seed)
    y = np.arange(height, dtype=np.float32)[:, None]
    x = np.arange(width, dtype=np.float32)[None, :]
    planes = []
    for c in range(channels):
        fx, fy = 0.002 + 0.0013 * c + 0.0007 * (seed % 7), 0.0017 + 0.0011 * c
        base = 128 + 70 * np.sin(2 * np.pi * (fx * x + 0.3 * c)) * np.cos(2 * np.pi * (fy * y + 0.2 * seed))
        tex = 25 * np.sin(2 * np.pi * (0.031 * x + 0.043 * y + c))
        noise = rng.integers(-6, 7, size=(height, width), dtype=np.int16)
        planes.append(np.clip(base + tex + noise, 0, 255).astype(np.uint8))
    img = np.stack(planes, axis=-1)
    return img[..., 0] if channels == 1 else img


def synth_frame(width, height, channels=3, seed=0):
    """Integer-only photo-like frame (exactly reproducible everywhere): two triangular-wave gradients
    per channel, a fine diagonal texture and +-4 of hashed noise. Compresses to ~0.2-0.3 B/px at q90 4:2:0."""
    y = np.arange(height, dtype=np.int64)[:, None]
    x = np.arange(width, dtype=np.int64)[None, :]

    def tri(t, period):
        t = t % period
        return np.abs(t - period // 2) * 510 // period  # 0..255

    planes = []
    for c in range(channels):
        a = tri(x * (3 + c) + y * (2 + seed % 5) + seed * 131 + c * 977, 2048 + 256 * c)
        b = tri(x * (1 + (seed >> 2) % 3) - y * (4 + c) + seed * 29, 1536 + 128 * (seed % 7))
        t = tri(x * 37 + y * 53 + c * 11, 64) >> 3
        hsh = (x * 0x9E3779B1 + y * 0x85EBCA77 + (seed * 4 + c) * 0xC2B2AE3D) & 0xFFFFFFFF
        hsh = (hsh ^ (hsh >> 15)) * 0x2C1B3C6D & 0xFFFFFFFF
        n = ((hsh >> 13) & 7) - 4
        planes.append(np.clip((a * 5 + b * 3) // 8 + t + n - 12, 0, 255).astype(np.uint8))
    img = np.stack(planes, axis=-1)
    return img[..., 0] if channels == 1 else img


def synth_frame_torch(width, height, channels=3, seed=0, row0=0, rows=None, device="cuda"):
    """synth_frame evaluated with torch integer ops on `device`, for rows [row0, row0 + rows): the
    same values as the numpy version (used for the 16384^2 workload, generated strip by strip)."""
    import torch
    rows = height - row0 if rows is None else rows
    y = torch.arange(row0, row0 + rows, dtype=torch.int64, device=device)[:, None]
    x = torch.arange(width, dtype=torch.int64, device=device)[None, :]

    def tri(t, period):
        t = t % period
        return (t - period // 2).abs() * 510 // period

    planes = []
    for c in range(channels):
        a = tri(x * (3 + c) + y * (2 + seed % 5) + seed * 131 + c * 977, 2048 + 256 * c)
        b = tri(x * (1 + (seed >> 2) % 3) - y * (4 + c) + seed * 29, 1536 + 128 * (seed % 7))
        t = tri(x * 37 + y * 53 + c * 11, 64) >> 3
        hsh = (x * 0x9E3779B1 + y * 0x85EBCA77 + (seed * 4 + c) * 0xC2B2AE3D) & 0xFFFFFFFF
        hsh = (hsh ^ (hsh >> 15)) * 0x2C1B3C6D & 0xFFFFFFFF
        n = ((hsh >> 13) & 7) - 4
        planes.append(((a * 5 + b * 3) // 8 + t + n - 12).clamp(0, 255).to(torch.uint8))
    img = torch.stack(planes, dim=-1)
    return img[..., 0].contiguous() if channels == 1 else img.contiguous()
