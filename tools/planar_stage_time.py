"""Stage times of the planar (ImageBuffer) path for one large image: warp kernel vs the generic kernel.

    python tools/planar_stage_time.py [size]          (GPU box)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jpeg_encoder_b200 as je  # noqa: E402


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    rng = np.random.default_rng(0)
    planes = [rng.integers(0, 256, size * size, dtype=np.uint8) for _ in range(3)]
    for sampling in ((2, 2), (1, 1), (4, 1)):
        for generic in ("0", "1"):
            os.environ["JPGB_FORCE_GENERIC_STAGE_A"] = generic
            enc = je.Encoder(90)
            enc.set_sampling_factor(je.SamplingFactor.from_factors(*sampling))
            dev = je.default_device(0)
            dev.set_timing(True)
            outs = [enc.encode_planes(planes, size, size, je.JpegColorType.Ycbcr) for _ in range(4)]
            t = dev.last_timing()
            dev.set_timing(False)
            print("planar ycbcr %dx%d sampling %s %s: colour_dct_quant %.3f ms (%d bytes)" %
                  (size, size, sampling, "generic" if generic == "1" else "warp", t["colour_dct_quant"], len(outs[-1])), flush=True)
    os.environ.pop("JPGB_FORCE_GENERIC_STAGE_A", None)


if __name__ == "__main__":
    main()
