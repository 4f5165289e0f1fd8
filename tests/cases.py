"""One description of an encoder configuration, applied both to the oracle and to the product Encoder."""
import jpeg_encoder_b200 as je
from oracle import oracle as orc

CT = {"luma": (orc.LUMA, je.ColorType.Luma), "rgb": (orc.RGB, je.ColorType.Rgb), "rgba": (orc.RGBA, je.ColorType.Rgba),
      "bgr": (orc.BGR, je.ColorType.Bgr), "bgra": (orc.BGRA, je.ColorType.Bgra), "ycbcr": (orc.YCBCR, je.ColorType.Ycbcr),
      "cmyk": (orc.CMYK, je.ColorType.Cmyk), "cmyk_as_ycck": (orc.CMYK_AS_YCCK, je.ColorType.CmykAsYcck),
      "ycck": (orc.YCCK, je.ColorType.Ycck)}
BPP = {"luma": 1, "rgb": 3, "rgba": 4, "bgr": 3, "bgra": 4, "ycbcr": 3, "cmyk": 4, "cmyk_as_ycck": 4, "ycck": 4}


def make_encoder(cfg, device=None):
    """cfg keys: quality, sampling (h, v), qtables, progressive_scans, restart_interval, optimize_huffman,
    density (unit, x, y), app_segments [(nr, bytes)]."""
    enc = je.Encoder(cfg.get("quality", 90), device=device)
    if cfg.get("sampling") is not None:
        enc.set_sampling_factor(je.SamplingFactor.from_factors(*cfg["sampling"]))
    if "qtables" in cfg:
        lu, ch = cfg["qtables"]
        enc.set_quantization_tables(je.QuantizationTableType(lu) if isinstance(lu, int) else lu,
                                    je.QuantizationTableType(ch) if isinstance(ch, int) else ch)
    if cfg.get("progressive_scans"):
        enc.set_progressive_scans(cfg["progressive_scans"])
    if cfg.get("restart_interval"):
        enc.set_restart_interval(cfg["restart_interval"])
    if cfg.get("optimize_huffman"):
        enc.set_optimized_huffman_tables(True)
    if "density" in cfg:
        u, x, y = cfg["density"]
        enc.set_density(je.PixelDensity((x, y), je.PixelDensityUnit(u)))
    for nr, data in cfg.get("app_segments", ()):
        enc.add_app_segment(nr, data)
    return enc


def oracle_encode(img, w, h, color, cfg):
    return orc.encode(img, w, h, CT[color][0], **cfg)


def gpu_encode(img, w, h, color, cfg, device=None):
    return make_encoder(cfg, device).encode(img, w, h, CT[color][1])
