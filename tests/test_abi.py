"""CPU-side checks: the C-ABI library builds, loads and exports every symbol include/jpegenc_b200.h
declares; host-side planning (no kernels) agrees with the oracle's geometry; the Python mirror
behaves like the reference's Encoder setters. No compute calls -- there is no GPU here."""
import ctypes as C
import os
import re

import pytest

import __graft_entry__ as entry
import jpeg_encoder_b200 as je

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_build_and_exported_symbols():
    lib_path = entry.build()
    assert os.path.exists(lib_path)
    header = open(os.path.join(ROOT, "include", "jpegenc_b200.h")).read()
    declared = set(re.findall(r"\b(jpgb_[a-z_0-9]+)\s*\(", header))
    declared -= {"jpgb_write_all_fn"}
    assert len(declared) >= 14
    lib = C.CDLL(lib_path)
    for name in sorted(declared):
        assert hasattr(lib, name), "library does not export %s" % name


def test_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", os.path.join(ROOT, "jpeg_encoder_b200", "libjpegenc_b200.so")],
                         capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(je.EncodingError) as e:
        je.Encoder(90).encode(bytes(12), 2, 2, je.ColorType.Rgb)
    assert e.value.kind == "Cuda"


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "jpeg_encoder_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.lower().replace("# the oracle", ""), "%s mentions the oracle" % f


def test_params_default_matches_encoder_new():
    lib = je.load_library()
    from jpeg_encoder_b200.encoder import _Params
    for q, samp in ((89, 0x22), (90, 0x11), (100, 0x11), (1, 0x22)):
        p = _Params()
        lib.jpgb_params_default(C.byref(p), q)
        assert (p.quality, p.sampling) == (q, samp)
        assert (p.density_unit, p.density_x, p.density_y) == (0, 1, 1)
        assert p.progressive_scans == 0 and p.restart_interval == 0 and p.optimize_huffman == 0
        assert je.Encoder(q).sampling_factor().get_sampling_factors() == ((samp >> 4), samp & 15)


def test_coef_layout_matches_reference_block_counts():
    """SURVEY.md 8a block counts: C1 32640+8160+8160, C2 512^2+2*256^2 (src/encoder.rs:713-717, 1012-1025)."""
    e = je.Encoder(90)
    e.set_sampling_factor(je.SamplingFactor.F_2_2)
    lay = e.coef_layout(1920, 1080, je.ColorType.Rgb)
    assert [lay.blocks_w[c] * lay.blocks_h[c] for c in range(3)] == [32640, 8160, 8160]
    assert lay.blocks_per_image == 48960 and lay.mcu_order == 1 and lay.blocks_per_mcu == 6  # interleaved: MCU order, padding included
    assert (lay.true_w[0], lay.true_h[0], lay.true_w[1], lay.true_h[1]) == (240, 135, 120, 68)
    lay = e.coef_layout(4096, 4096, je.ColorType.Rgb)
    assert lay.blocks_per_image == 512 * 512 + 2 * 256 * 256
    e.set_sampling_factor(je.SamplingFactor.F_4_1)
    lay = e.coef_layout(258, 128, je.ColorType.Rgb)  # 9 MCU columns of 32 px
    assert (lay.blocks_w[0], lay.blocks_h[0], lay.true_w[0], lay.true_w[1]) == (36, 16, 33, 9)
    e = je.Encoder(90)
    e.set_sampling_factor(je.SamplingFactor.F_2_2)
    e.set_optimized_huffman_tables(True)  # sequential: the true grids encode_blocks walks, raster order, no padding blocks
    lay = e.coef_layout(1920, 1080, je.ColorType.Rgb)
    assert lay.mcu_order == 0 and lay.blocks_per_image == 240 * 135 + 2 * 120 * 68
    assert list(lay.block_offset)[:3] == [0, 240 * 135, 240 * 135 + 120 * 68]
    lay = je.Encoder(95).coef_layout(8192, 8192, je.ColorType.CmykAsYcck)
    assert lay.n_components == 4 and lay.blocks_per_image == 4 * 1024 * 1024


def test_setters_mirror_reference():
    e = je.Encoder(100)
    e.set_progressive(True)
    assert e.progressive_scans() == 4  # src/encoder.rs:1323-1331
    e.set_progressive(False)
    assert e.progressive_scans() is None
    with pytest.raises(ValueError):
        e.set_progressive_scans(1)
    with pytest.raises(ValueError):
        e.set_progressive_scans(65)
    e.set_restart_interval(0)
    assert e.restart_interval() is None
    with pytest.raises(je.EncodingError) as err:
        e.add_app_segment(0, b"x")
    assert err.value.kind == "InvalidAppSegment"
    with pytest.raises(je.EncodingError) as err:
        e.add_app_segment(1, bytes(65534))
    assert err.value.kind == "AppSegmentTooLarge"
    with pytest.raises(je.EncodingError) as err:
        e.add_icc_profile(bytes(255 * 65519))
    assert err.value.kind == "IccTooLarge"
    assert je.PixelDensity.dpi(300) == je.PixelDensity((300, 300), je.PixelDensityUnit.Inches)
    for f in je.SamplingFactor:  # src/encoder.rs:1302-1321
        h, v = f.get_sampling_factors()
        assert je.SamplingFactor.from_factors(h, v).get_sampling_factors() == (h, v)
    assert je.SamplingFactor.R_4_2_0.get_sampling_factors() == (2, 2)
    assert je.SamplingFactor.from_factors(4, 4) is None


def test_cpp_mirror_compiles_and_fails_loudly_without_gpu(tmp_path):
    """include/jpeg_encoder.hpp (the compiled-language mirror of the reference API) builds against the C ABI."""
    import subprocess
    import torch
    exe = str(tmp_path / "mirror_smoke")
    pkg = os.path.join(ROOT, "jpeg_encoder_b200")
    subprocess.check_call(["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "mirror_smoke.cpp"),
                           "-o", exe, "-L" + pkg, "-ljpegenc_b200", "-Wl,-rpath," + pkg])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    if not torch.cuda.is_available():
        assert "EncodingError 8" in r.stdout  # JPGB_ERR_CUDA: no fallback


# ---- host planner against the oracle, no GPU -----------------------------------------------------
def _first_scan_data_offset(jpg):
    """offset just past the first SOS segment"""
    i = 2
    while True:
        assert jpg[i] == 0xFF
        m = jpg[i + 1]
        n = int.from_bytes(jpg[i + 2:i + 4], "big")
        i += 2 + n
        if m == 0xDA:
            return i


@pytest.mark.parametrize("color", ["luma", "rgb", "bgra", "ycbcr", "cmyk", "cmyk_as_ycck", "ycck"])
def test_header_bytes_match_oracle(color):
    """Container segments written by the C++ planner (csrc/host.cpp) == the oracle's, for every mode that
    uses the default Huffman tables (Q10 truncated DQT, Q11 component ids, Q21 segment order, DRI, APPn)."""
    import numpy as np
    from cases import BPP, CT, make_encoder, oracle_encode
    rng = np.random.default_rng(1)
    big = rng.integers(1, 3000, 64).tolist()
    cfgs = [dict(quality=90, sampling=(2, 2)), dict(quality=30, sampling=(4, 1), restart_interval=77),
            dict(quality=100, sampling=(1, 2), progressive_scans=7), dict(quality=55, qtables=(big, 3), sampling=(2, 4)),
            dict(quality=75, density=(1, 300, 72), app_segments=[(1, b"Exif\0\0abc"), (15, bytes(range(200)))], progressive_scans=2,
                 restart_interval=1)]
    w, h = 37, 21
    img = rng.integers(0, 256, (h, w, BPP[color]), dtype=np.uint8)
    for cfg in cfgs:
        want = oracle_encode(img if BPP[color] > 1 else img[..., 0], w, h, color, cfg)
        got = make_encoder(cfg).build_header(w, h, CT[color][1])
        assert got == want[:_first_scan_data_offset(want)], cfg


def test_optimized_huffman_host_matches_oracle():
    """Annex K.2 as run by the planner == the oracle's restatement, on skewed, flat and sparse histograms."""
    import numpy as np
    from oracle import oracle as orc
    lib = je.load_library()
    rng = np.random.default_rng(2)
    cases = []
    for _ in range(40):
        f = np.zeros(257, np.uint32)
        k = int(rng.integers(1, 257))
        idx = rng.choice(256, k, replace=False)
        f[idx] = (rng.pareto(0.7, k) * 10 + 1).astype(np.uint32) if rng.random() < 0.7 else rng.integers(1, 50, k)
        f[256] = 1
        cases.append(f)
    fib = np.zeros(257, np.uint32)  # Fibonacci-like frequencies force the > 16 bit length limiting of Figure K.3
    a, b = 1, 1
    for i in range(30):
        fib[i] = a
        a, b = b, a + b
    fib[256] = 1
    cases.append(fib)
    for f in cases:
        length = (C.c_uint8 * 16)()
        values = (C.c_uint8 * 256)()
        n = C.c_uint32()
        rc = lib.jpgb_optimized_huffman_table(f.ctypes.data_as(C.POINTER(C.c_uint32)), length, values, C.byref(n))
        assert rc == 0
        want_len, want_vals = orc.huffman_optimized(f.tolist())
        assert list(length) == want_len and list(values)[:n.value] == want_vals


def test_host_entry_points_survive_arbitrary_params():
    """The planner behind the host-only entry points takes whatever bytes a caller puts into jpgb_params (pointers
    excepted): it must answer with JPGB_OK / BAD_PARAMS / ZERO_DIMENSIONS consistently, never crash or disagree."""
    import random
    from jpeg_encoder_b200.encoder import _CoefLayout, _Params, _Strip
    lib = je.load_library()
    rnd = random.Random(4)
    seen = set()
    for _ in range(6000):
        p = _Params()
        raw = (C.c_uint8 * C.sizeof(p)).from_buffer(p)
        for i in range(C.sizeof(p)):
            raw[i] = rnd.getrandbits(8) if rnd.random() < 0.5 else rnd.choice([0, 1, 2, 4, 0x11, 0x22, 0x41, 9, 255])
        p.n_app, p.apps = 0, None  # pointers must be valid by contract
        if rnd.random() < 0.7:
            p.color_type = rnd.randrange(0, 11)
            p.sampling = rnd.choice([0x11, 0x12, 0x21, 0x22, 0x41, 0x42, 0x14, 0x24, 0x44, 0x33, 0x00, 0x91, 0xA2])
            p.progressive_scans = rnd.choice([0, 1, 2, 4, 64, 65, 255])
            p.qtable_kind[0], p.qtable_kind[1] = rnd.randrange(0, 12), rnd.randrange(0, 12)
        n, ns, lay, k = C.c_size_t(), C.c_uint32(), _CoefLayout(), C.c_uint32()
        arr = (_Strip * 8)()
        r1 = lib.jpgb_build_header(C.byref(p), None, 0, C.byref(n))
        r2 = lib.jpgb_scan_count(C.byref(p), C.byref(ns))
        r3 = lib.jpgb_coef_layout_for(C.byref(p), C.byref(lay))
        r4 = lib.jpgb_plan_strips(C.byref(p), 8, arr, C.byref(k))
        assert r1 == r2 == r3 and r1 in (0, 2, 5)          # one verdict on the settings
        assert r4 == r1 or (r1 == 0 and r4 == 5)           # strips additionally need a restart interval
        if r1 == 0:
            buf = (C.c_uint8 * n.value)()
            assert lib.jpgb_build_header(C.byref(p), buf, n.value, C.byref(n)) == 0 and bytes(buf[:2]) == b"\xff\xd8"
            assert 1 <= ns.value <= 4 * 64 and lay.blocks_per_image > 0
        seen.add(r1)
    assert seen == {0, 2, 5}
