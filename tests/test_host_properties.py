"""Property tests (hypothesis) of the host-side logic that needs no GPU: the C++ planner's container bytes against the
oracle for random settings, and the strip planner's invariants against an independent restatement."""
import numpy as np
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402

import jpeg_encoder_b200 as je  # noqa: E402
from cases import BPP, CT, make_encoder, oracle_encode  # noqa: E402

SAMPLINGS = [(1, 1), (2, 1), (1, 2), (2, 2), (4, 1), (4, 2), (1, 4), (2, 4)]
NCOMP = {"luma": 1, "rgb": 3, "rgba": 3, "bgr": 3, "bgra": 3, "ycbcr": 3, "cmyk": 4, "cmyk_as_ycck": 4, "ycck": 4}


def _first_scan_data_offset(jpg):
    i = jpg.index(b"\xff\xda")
    return i + 2 + int.from_bytes(jpg[i + 2:i + 4], "big")


@st.composite
def _settings(draw):
    cfg = dict(quality=draw(st.integers(0, 255)), sampling=draw(st.sampled_from(SAMPLINGS)))
    if draw(st.booleans()):
        cfg["progressive_scans"] = draw(st.integers(2, 64))
    if draw(st.booleans()):
        cfg["restart_interval"] = draw(st.integers(1, 65535))
    kind = draw(st.integers(0, 2))
    if kind == 1:
        cfg["qtables"] = (draw(st.integers(0, 8)), draw(st.integers(0, 8)))
    elif kind == 2:
        cfg["qtables"] = (draw(st.lists(st.integers(0, 65535), min_size=64, max_size=64)), draw(st.integers(0, 8)))
    if draw(st.booleans()):
        cfg["density"] = (draw(st.integers(0, 2)), draw(st.integers(0, 65535)), draw(st.integers(0, 65535)))
    if draw(st.booleans()):
        cfg["app_segments"] = [(draw(st.integers(1, 15)), draw(st.binary(max_size=300))) for _ in range(draw(st.integers(1, 3)))]
    return cfg


@settings(max_examples=150, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(st.sampled_from(sorted(BPP)), st.integers(1, 300), st.integers(1, 300), _settings())
def test_planner_header_equals_oracle_header(color, w, h, cfg):
    """SOI .. first SOS as csrc/host.cpp writes them == the oracle's file prefix (default Huffman tables: the planner's
    header does not depend on the pixels), for any settings: Q8 quality clamp, Q10 truncated DQT, Q11 ids, Q12 sampling
    per colour type, Q13 mode, Q21 order, DRI, density, APPn."""
    img = np.zeros((h, w, BPP[color]), np.uint8)
    want = oracle_encode(img if BPP[color] > 1 else img[..., 0], w, h, color, cfg)
    got = make_encoder(cfg).build_header(w, h, CT[color][1])
    assert got == want[:_first_scan_data_offset(want)]


def _units_per_mcu_row(color, w, sampling, progressive, optimize):
    """Restart units per MCU row of every scan, restated from SURVEY.md Q12-Q15 (not from the planner)."""
    hs, vs = sampling
    if color == "luma":
        comps = [(1, 1)]
    elif NCOMP[color] == 3:
        comps = [(hs, vs), (1, 1), (1, 1)]
    elif color == "cmyk":
        comps = [(1, 1), (1, 1), (1, 1), (hs, vs)]
    else:
        comps = [(hs, vs), (1, 1), (1, 1), (hs, vs)]
    hmax, vmax = max(c[0] for c in comps), max(c[1] for c in comps)
    interleaved = not progressive and not optimize and 4 not in sampling  # mode follows the *setting* (Q13)
    if interleaved:
        return [-(-w // (8 * hmax))], vmax
    out = []
    for ch, cv in comps:
        tw = -(-(-(-w // 8)) // (hmax // ch))
        out.append(cv * tw)  # a component's blocks in one MCU row: V_c block rows of its true grid
    return out, vmax


@settings(max_examples=300, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(st.sampled_from(sorted(BPP)), st.integers(1, 4000), st.integers(1, 4000), st.sampled_from(SAMPLINGS), st.integers(1, 3000),
       st.booleans(), st.booleans(), st.integers(1, 8))
def test_strip_plan_invariants(color, w, h, sampling, restart, progressive, optimize, max_strips):
    enc = je.Encoder(80)
    enc.set_sampling_factor(je.SamplingFactor.from_factors(*sampling))
    enc.set_restart_interval(restart)
    if progressive:
        enc.set_progressive_scans(4)
    enc.set_optimized_huffman_tables(optimize)
    strips = enc.plan_strips(w, h, CT[color][1], max_strips)
    units, vmax = _units_per_mcu_row(color, w, sampling, progressive, optimize)
    assert 1 <= len(strips) <= max_strips
    assert strips[0][0] == 0 and sum(n for _, n in strips) == h
    assert all(a[0] + a[1] == b[0] for a, b in zip(strips, strips[1:]))  # contiguous, in order
    for r0, rows in strips:
        assert rows > 0 and r0 % (8 * vmax) == 0
        mcu_row0 = r0 // (8 * vmax)
        for u in units:  # every scan: the strip starts on a restart boundary
            assert (mcu_row0 * u) % restart == 0
