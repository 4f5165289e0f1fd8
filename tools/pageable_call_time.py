"""Time of the drop-in call with a pageable source (what Encoder::encode(&[u8]) hands over) for a few image sizes.

    python tools/pageable_call_time.py            (GPU box; JPGB_COPY_THREADS=1 = staging copy by the caller alone)
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jpeg_encoder_b200 as je  # noqa: E402


def main():
    rng = np.random.default_rng(1)
    for w, h in ((1920, 1080), (4096, 4096), (8192, 8192)):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        enc = je.Encoder(90)
        enc.set_sampling_factor(je.SamplingFactor.F_2_2)
        for _ in range(4):
            enc.encode(img, w, h, je.ColorType.Rgb)
        n = 40 if w < 4096 else 8
        t = time.perf_counter()
        for _ in range(n):
            enc.encode(img, w, h, je.ColorType.Rgb)
        dt = (time.perf_counter() - t) / n
        print("%dx%d pageable: %.3f ms per call, %.0f MP/s (JPGB_COPY_THREADS=%s)" % (w, h, dt * 1e3, w * h / dt / 1e6, os.environ.get("JPGB_COPY_THREADS", "unset")), flush=True)


if __name__ == "__main__":
    main()
