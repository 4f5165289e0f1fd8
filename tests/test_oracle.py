"""Pins the CPU oracle (oracle/jpeg_oracle.c) against every known-answer vector the reference's own
tests hold for the encode path and re-creates its 21 round-trip tests
(/root/reference/src/lib.rs:188-553) with PIL as the decoder and the same |diff| < 20 tolerance."""
import io
import json
import os

import numpy as np
import pytest
from PIL import Image

import images
from oracle import oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden")


# ---- committed file hashes (tests/golden/make_golden.py) ---------------------------------------
def _golden():
    import sys
    sys.path.insert(0, GOLD)
    import make_golden
    return make_golden, json.load(open(os.path.join(GOLD, "oracle_jpegs.json")))


def test_oracle_reproduces_committed_file_hashes():
    """The oracle's file bytes for the fixed configuration matrix have not drifted since they were committed."""
    import hashlib
    mg, want = _golden()
    cases = mg.golden_matrix()
    assert sorted(c[0] for c in cases) == sorted(want)
    for case in cases:
        data = mg.encode_case(case)
        assert (hashlib.sha256(data).hexdigest(), len(data)) == (want[case[0]]["sha256"], want[case[0]]["len"]), case[0]


# ---- unit known-answer tests ------------------------------------------------------------------
def test_rgb_to_ycbcr_kat():
    """src/image_buffer.rs:325-422 (5 primaries + 88 libjpeg-derived triples)."""
    trip = json.load(open(os.path.join(GOLD, "rgb_to_ycbcr_kat.json")))
    assert len(trip) == 93
    for r, g, b, y, cb, cr in trip:
        assert orc.rgb_to_ycbcr(r, g, b) == (y, cb, cr)


def test_fdct_kat():
    """src/fdct.rs:249-285 (libjpeg jpeg_fdct_islow vectors)."""
    d = json.load(open(os.path.join(GOLD, "fdct_kat.json")))
    assert orc.fdct(d["INPUT1"]).tolist() == d["OUTPUT1"]
    assert orc.fdct(d["INPUT2"]).tolist() == d["OUTPUT2"]
    assert orc.fdct(d["INPUT1"], i16model=True).tolist() == d["OUTPUT1"]
    assert orc.fdct(d["INPUT2"], i16model=True).tolist() == d["OUTPUT2"]


def test_fdct_avx2_model_equals_scalar():
    """The reference never asserts fdct_avx2 == fdct; the `simd` feature is what Encoder::encode runs
    on x86. Check the 16-bit-stage model of src/avx2/fdct.rs against the scalar path on extremes,
    +-1 checkerboards and random level-shifted 8-bit blocks."""
    rng = np.random.default_rng(1)
    blocks = [np.full(64, -128), np.full(64, 127), np.tile([-128, 127], 32), np.tile([127, -128], 32)]
    cb = np.indices((8, 8)).sum(axis=0) % 2
    blocks += [np.where(cb, 127, -128).reshape(-1), np.where(cb, -128, 127).reshape(-1)]
    for k in range(8):  # cosine-like extremal patterns per frequency
        u = np.cos((2 * np.arange(8) + 1) * k * np.pi / 16)
        for l in range(8):
            v = np.cos((2 * np.arange(8) + 1) * l * np.pi / 16)
            blocks.append(np.where(np.outer(u, v) >= 0, 127, -128).reshape(-1))
    blocks += [rng.integers(-128, 128, 64) for _ in range(3000)]
    for b in blocks:
        assert orc.fdct(b).tolist() == orc.fdct(b, i16model=True).tolist()


def test_simd_baseline_path_is_bit_identical():
    """The AVX2 colour conversion and fDCT that bench.py's CPU legs switch on (the role of the crate's `simd`
    feature, src/avx2/) must not change a single bit: blocks, extremes (src/avx2/ycbcr.rs:191-237 tests the same
    equality for colour) and whole files."""
    if not orc.has_simd():
        pytest.skip("oracle built without AVX2")
    rng = np.random.default_rng(7)
    for _ in range(500):
        b = rng.integers(-128, 128, 64).astype(np.int16)
        assert (orc.fdct(b) == orc.fdct(b, simd=True)).all()
    for b in (np.full(64, -128, np.int16), np.full(64, 127, np.int16), np.tile(np.array([-128, 127], np.int16), 32)):
        assert (orc.fdct(b) == orc.fdct(b, simd=True)).all()
    for color, w, h, kw in (("rgb", 259, 67, dict(quality=90, sampling=(2, 2))), ("rgba", 64, 33, dict(quality=75)),
                            ("rgb", 9, 9, dict(quality=100, sampling=(1, 1), progressive_scans=4)),
                            ("bgr", 40, 40, dict(quality=80, optimize_huffman=True))):
        ct = {"rgb": orc.RGB, "rgba": orc.RGBA, "bgr": orc.BGR}[color]
        img = rng.integers(0, 256, (h, w, orc.BPP[ct]), dtype=np.uint8)
        try:
            orc.set_simd(False)
            a = orc.encode(img, w, h, ct, **kw)
            orc.set_simd(True)
            b = orc.encode(img, w, h, ct, **kw)
        finally:
            orc.set_simd(False)
        assert a == b


def test_quant_new_100():
    """src/quantization.rs:314-338."""
    for luma in (True, False):
        tab, rec, cor = orc.quant_table(0, 100, luma)
        assert all(v == 1 << 3 for v in tab)
    tab, rec, cor = orc.quant_table(0, 100, True)
    for i in range(-255, 255):
        assert orc.quantize(i << 3, rec[0], cor[0]) == i


def test_quant_spot_values():
    """SURVEY.md Q7/Q8 spot values (derived from src/quantization.rs:187-283)."""
    tab, rec, cor = orc.quant_table(0, 90, True)
    assert [v >> 3 for v in tab[:8]] == [3, 2, 2, 3, 5, 8, 10, 12]
    tab, rec, cor = orc.quant_table(0, 85, True)
    assert [v >> 3 for v in tab[:8]] == [5, 3, 3, 5, 7, 12, 15, 18]
    t2, r2, c2 = orc.quant_table(9, 0, True, custom=[1, 3, 5, 255, 5000] + [1] * 59)
    assert t2[:5] == [8, 24, 40, 2040, 2048 << 3]
    assert (r2[0], c2[0]) == (4096, 4) and (r2[1], c2[1]) == (1365, 13)
    assert (r2[2], c2[2]) == (819, 21) and (r2[3], c2[3]) == (16, 1021)
    assert orc.quantize(4092, r2[1], c2[1]) == 170  # differs from round-half-up division (171)


def test_get_num_bits_matches_get_code():
    """src/encoder.rs:1286-1300."""
    for v in range(-(2 ** 13), 2 ** 13 + 1):
        assert orc.get_num_bits(v) == orc.get_code(v)[0]
    assert orc.get_code(-1) == (1, 0) and orc.get_code(1) == (1, 1)
    assert orc.get_code(-3) == (2, 0) and orc.get_code(5) == (3, 5) and orc.get_code(-5) == (3, 2)


def test_huffman_optimized_small():
    """K.2 on a tiny alphabet: two real symbols + the reserved one (src/huffman.rs:99-221)."""
    freq = [0] * 257
    freq[0], freq[1], freq[256] = 10, 3, 1
    length, values = orc.huffman_optimized(freq)
    assert sum(length) == 2 and sorted(values) == [0, 1]
    assert length[0] == 1 and length[1] == 1 and values == [0, 1]


# ---- stream-level known answers ---------------------------------------------------------------
def _scan_payload(jpg):
    i = jpg.rindex(b"\xFF\xDA")
    n = int.from_bytes(jpg[i + 2:i + 4], "big")
    assert jpg[-2:] == b"\xFF\xD9"
    return jpg[i + 2 + n:-2]


def test_flat_gray_streams():
    """Hand-derived payloads: flat mid-grey => all coefficients zero."""
    j = orc.encode(np.full((8, 8), 128, np.uint8), 8, 8, orc.LUMA, quality=90)
    assert _scan_payload(j) == bytes([0x2B])
    j = orc.encode(np.full((8, 8, 3), 128, np.uint8), 8, 8, orc.RGB, quality=90, sampling=(1, 1))
    assert _scan_payload(j) == bytes([0x28, 0x03])
    j = orc.encode(np.full((16, 16, 3), 128, np.uint8), 16, 16, orc.RGB, quality=90, sampling=(2, 2))
    assert _scan_payload(j) == bytes([0x28, 0xA2, 0x8A, 0x00])


def test_header_known_answer_c1():
    """Header layout for BASELINE config 1 (1920x1080 RGB q90 F_2_2), SURVEY.md section 0."""
    img = images.bench_img(1920, 1080)
    j = orc.encode(img, 1920, 1080, orc.RGB, quality=90, sampling=(2, 2))
    exp = bytes.fromhex("FFD8" "FFE000104A46494600010200000100010000")
    assert j.startswith(exp)
    o = len(exp)
    assert j[o:o + 19] == bytes.fromhex("FFC0001108" "0438" "0780" "03" "002200" "011101" "021101")
    o += 19
    assert j[o:o + 5] == bytes.fromhex("FFDB004300")
    tab, _, _ = orc.quant_table(0, 90, True)
    zz = [0, 1, 8, 16, 9, 2, 3, 10]
    assert list(j[o + 5:o + 13]) == [tab[z] >> 3 for z in zz]
    o += 69
    assert j[o:o + 5] == bytes.fromhex("FFDB004301")
    o += 69
    assert j[o:o + 5] == bytes.fromhex("FFC4001F00")
    o += 2 + 0x1F
    assert j[o:o + 5] == bytes.fromhex("FFC400B510")
    o += 2 + 0xB5
    assert j[o:o + 5] == bytes.fromhex("FFC4001F01")
    o += 2 + 0x1F
    assert j[o:o + 5] == bytes.fromhex("FFC400B511")
    o += 2 + 0xB5
    assert j[o:o + 14] == bytes.fromhex("FFDA000C03" "0000" "0111" "0211" "003F00")
    assert j[-2:] == b"\xFF\xD9"


# ---- the reference's 21 round-trip tests, PIL as decoder ----------------------------------------
def _decode(jpg):
    im = Image.open(io.BytesIO(jpg))
    im.load()
    return im


def _check(data, jpg, mode):
    im = _decode(jpg)
    assert im.mode == mode
    assert im.size == (data.shape[1], data.shape[0])
    dec = np.asarray(im).astype(np.int16)
    ref = data.astype(np.int16)
    # CMYK: the file stores 255-c (Q3); PIL un-inverts Adobe CMYK/YCCK itself, so dec is comparable as is
    assert dec.shape == ref.shape
    assert np.abs(dec - ref).max() < 20, np.abs(dec - ref).max()


RGB_CASES = {
    "rgb_100": dict(quality=100),
    "rgb_80": dict(quality=80),
    "custom_q_table": dict(quality=100, qtables=([1] * 64, [1] * 64)),
    "2_2": dict(quality=100, sampling=(2, 2)),
    "2_1": dict(quality=100, sampling=(2, 1)),
    "4_1": dict(quality=100, sampling=(4, 1)),
    "1_1": dict(quality=100, sampling=(1, 1)),
    "1_4": dict(quality=100, sampling=(1, 4)),
    "progressive": dict(quality=100, sampling=(2, 1), progressive_scans=4),
    "optimized": dict(quality=100, sampling=(2, 2), optimize_huffman=True),
    "optimized_progressive": dict(quality=100, sampling=(2, 1), progressive_scans=4, optimize_huffman=True),
    "restart": dict(quality=100, restart_interval=32),
    "restart_4_1": dict(quality=100, sampling=(4, 1), restart_interval=32),
    "restart_progressive": dict(quality=85, progressive_scans=4, restart_interval=32),
}


@pytest.mark.parametrize("name", sorted(RGB_CASES))
def test_roundtrip_rgb(name):
    """src/lib.rs:201-381, 409-481."""
    kw = RGB_CASES[name]
    img = images.ref_img_rgb()
    jpg = orc.encode(img, 258, 128, orc.RGB, **kw)
    if "restart" in name:
        assert b"\xFF\xDD\x00\x04\x00\x20" in jpg  # DRI bytes, src/lib.rs:417
    _check(img, jpg, "RGB")


def test_roundtrip_gray_100():
    img = images.ref_img_gray()
    _check(img, orc.encode(img, 258, 128, orc.LUMA, quality=100), "L")


def test_roundtrip_rgba_80():
    jpg = orc.encode(images.ref_img_rgba(), 258, 128, orc.RGBA, quality=80)
    _check(images.ref_img_rgb(), jpg, "RGB")
    assert jpg == orc.encode(images.ref_img_rgb(), 258, 128, orc.RGB, quality=80)


def test_roundtrip_cmyk_and_ycck():
    """src/lib.rs:383-407."""
    img = images.ref_img_cmyk()
    j1 = orc.encode(img, 258, 192, orc.CMYK, quality=100)
    assert b"Adobe\0\0\0\0\0\0\0" in j1
    _check(img, j1, "CMYK")
    j2 = orc.encode(img, 258, 192, orc.CMYK_AS_YCCK, quality=100)
    assert b"Adobe\0\0\0\0\0\0\x02" in j2
    _check(img, j2, "CMYK")


def test_app_segment_bytes():
    """src/lib.rs:483-504."""
    jpg = orc.encode(images.ref_img_rgb(), 258, 128, orc.RGB, quality=100, app_segments=[(15, b"HOHOHO\0")])
    assert b"\xEF\x00\x09HOHOHO\x00" in jpg


def test_icc_profile_roundtrip():
    """src/lib.rs:506-539; chunking as add_icc_profile, src/encoder.rs:392-417."""
    icc = bytes(i % 255 for i in range(128 * 1024))
    maxc = 65535 - 2 - 12 - 2
    chunks = [icc[i:i + maxc] for i in range(0, len(icc), maxc)]
    segs = [(2, b"ICC_PROFILE\0" + bytes([i + 1, len(chunks)]) + c) for i, c in enumerate(chunks)]
    jpg = orc.encode(images.ref_img_rgb(), 258, 128, orc.RGB, quality=100, app_segments=segs)
    assert b"ICC_PROFILE\0" in jpg
    assert _decode(jpg).info.get("icc_profile") == icc


def test_optimized_missing_table_frequency_1x1():
    """src/lib.rs:541-553 (1x1, F_2_2, optimized)."""
    data = np.array([[[0xFB, 0x15, 0x15]]], np.uint8)
    jpg = orc.encode(data, 1, 1, orc.RGB, quality=100, sampling=(2, 2), optimize_huffman=True)
    _check(data, jpg, "RGB")


# ---- error behaviour (src/encoder.rs:447-454, 521-526) -------------------------------------------
def test_errors():
    with pytest.raises(orc.OracleError) as e:
        orc.encode(np.zeros(10, np.uint8), 4, 4, orc.RGB)
    assert e.value.code == 1
    with pytest.raises(orc.OracleError) as e:
        orc.encode(np.zeros(10, np.uint8), 0, 4, orc.RGB)
    assert e.value.code == 2
    # extra trailing bytes are ignored (Q22)
    img = images.ref_img_rgb(16, 16)
    a = orc.encode(img, 16, 16, orc.RGB)
    b = orc.encode(np.concatenate([img.reshape(-1), np.arange(100, dtype=np.uint8)]), 16, 16, orc.RGB)
    assert a == b


# ---- decode sanity over the wider matrix (files must be valid JPEGs) ---------------------------
@pytest.mark.parametrize("sampling", [(1, 1), (1, 2), (2, 1), (2, 2), (4, 1), (4, 2), (1, 4), (2, 4)])
@pytest.mark.parametrize("mode", ["baseline", "optimized", "progressive"])
def test_decodes_all_samplings(sampling, mode):
    img = images.photo_like(75, 53, 3, seed=3)
    kw = dict(quality=85, sampling=sampling)
    if mode == "optimized":
        kw["optimize_huffman"] = True
    if mode == "progressive":
        kw["progressive_scans"] = 5
    jpg = orc.encode(img, 75, 53, orc.RGB, **kw)
    im = _decode(jpg)
    assert im.size == (75, 53)
    dec = np.asarray(im.convert("L")).astype(np.int16)
    ref = np.asarray(Image.fromarray(img).convert("L")).astype(np.int16)
    assert np.abs(dec - ref).mean() < 6
