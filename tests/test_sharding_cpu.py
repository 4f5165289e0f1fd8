"""Host-side logic of the multi-GPU paths on CPU: batch sharding, strip planning through the C ABI
(no kernels), scan-major assembly, and the world_size-2 gather over gloo."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import jpeg_encoder_b200 as je
from jpeg_encoder_b200 import sharding


def test_shard_batch_covers_everything():
    for n in (1, 7, 8, 1024, 1025):
        for world in (1, 2, 4, 8):
            spans = [sharding.shard_batch(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _enc(restart, sampling=je.SamplingFactor.F_2_2, progressive=None, optimize=False):
    e = je.Encoder(90)
    e.set_sampling_factor(sampling)
    e.set_restart_interval(restart)
    if progressive:
        e.set_progressive_scans(progressive)
    e.set_optimized_huffman_tables(optimize)
    return e


def test_plan_strips_config5_geometry():
    """16384^2 RGB progressive 4:2:0, restart 2048: 8 strips of 128 MCU rows."""
    strips = _enc(2048, progressive=4).plan_strips(16384, 16384, je.ColorType.Rgb, 8)
    assert strips == [(i * 2048, 2048) for i in range(8)]
    assert _enc(2048, progressive=4).scan_count(16384, 16384, je.ColorType.Rgb) == 12


def test_plan_strips_alignment_rules():
    # luma 240 blocks per row, chroma 120: restart 64 needs groups of 8 MCU rows (8*120 % 64 == 0)
    strips = _enc(64, progressive=4).plan_strips(1920, 1080, je.ColorType.Rgb, 8)
    assert all(r % (8 * 16) == 0 for r, _ in strips)
    assert sum(n for _, n in strips) == 1080 and strips[0][0] == 0
    # interleaved: units are MCUs (120 per row): restart 7 only aligns every 7 MCU rows
    strips = _enc(7).plan_strips(1920, 1080, je.ColorType.Rgb, 4)
    assert all(r % (7 * 16) == 0 for r, _ in strips)
    assert sum(n for _, n in strips) == 1080
    # never more strips than aligned groups; a single group gives one strip
    assert len(_enc(65535).plan_strips(640, 480, je.ColorType.Rgb, 8)) == 1
    with pytest.raises(je.EncodingError):
        _enc(0).plan_strips(1920, 1080, je.ColorType.Rgb, 8)  # no restart interval: replicas only
    # optimized tables: sequential scans (one per component); the strips exchange histograms first
    strips = _enc(64, optimize=True).plan_strips(1920, 1080, je.ColorType.Rgb, 8)
    assert len(strips) > 1 and sum(n for _, n in strips) == 1080


def test_merge_strip_histograms_rechains_the_first_dc():
    """The reference's histogram chains DC differences over the whole component without restart resets
    (src/encoder.rs:1086-1200, Q17): a strip counted its first block against 0, the merge re-chains it."""
    e = _enc(64, optimize=True)
    W = je.encoder.HIST_WORDS
    hist = [0] * W
    # two strips; table 0 <- Y, table 1 <- Cb + Cr. Strip 1: first DCs (Y, Cb, Cr) = (100, -3, 0); strip 0 last DCs = (90, 5, 0)
    # each strip contributed cat(first - 0): strip 0 firsts are (7, 0, 0) -> cats 3, 0, 0; strip 1 -> cats 7, 2, 0
    hist[0 * 257 + 3] += 1
    hist[0 * 257 + 7] += 1
    hist[2 * 257 + 0] += 3
    hist[2 * 257 + 2] += 1
    hist[1 * 257 + 0x11] = 5  # AC bins are left alone
    edge = [7, 0, 0, 0, 90, 5, 0, 0,   100, -3, 0, 0, 1, 1, 1, 0]
    out = e.merge_strip_histograms(hist, edge, 1920, 1080, je.ColorType.Rgb)
    want = list(hist)
    want[0 * 257 + 7] -= 1; want[0 * 257 + 4] += 1      # Y: 100 - 90 = 10 -> category 4
    want[2 * 257 + 2] -= 1; want[2 * 257 + 4] += 1      # Cb: -3 - 5 = -8 -> category 4
    # Cr: 0 - 0 = 0 -> category 0, unchanged
    assert out == want


def test_assemble_is_scan_major():
    a = [b"H0", b"S1a", b"S2a"]
    b = [b"r0", b"r1", b"r2E"]
    assert sharding.assemble_pieces([a, b]) == b"H0r0S1ar1S2ar2E"
    assert sharding.split_pieces(b"abcdef", [0, 2, 2, 6]) == [b"ab", b"", b"cdef"]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank r owns pieces of 3 scans with rank-dependent sizes
    pieces = [bytes([10 * rank + k]) * (3 + 2 * rank + k) for k in range(3)]
    offs = [0]
    for pc in pieces:
        offs.append(offs[-1] + len(pc))
    buf = torch.tensor(list(b"".join(pieces)), dtype=torch.uint8)
    out = sharding.gather_strip_pieces(buf, offs, rank, world, torch.device("cpu"))
    if rank == 0:
        q.put(bytes(out.tolist()))
    # optimized tables with strips: histogram all-reduce + edge all-gather
    hist = [rank + 1] * je.encoder.HIST_WORDS
    hs, edges = sharding.exchange_strip_histograms(hist, [rank * 10 + i for i in range(8)], torch.device("cpu"))
    assert hs == [sum(r + 1 for r in range(world))] * je.encoder.HIST_WORDS
    assert edges == [r * 10 + i for r in range(world) for i in range(8)]
    lo, hi = sharding.shard_batch(11, world, rank)
    t = torch.tensor([hi - lo])
    dist.all_reduce(t)
    assert int(t) == 11
    dist.destroy_process_group()


def test_gather_strip_pieces_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    exp = sharding.assemble_pieces([[bytes([10 * r + k]) * (3 + 2 * r + k) for k in range(3)] for r in range(world)])
    assert got == exp


def test_piece_destinations_is_the_scan_major_order():
    """The placement arithmetic of the device gather (csrc/gather.cu, restated in sharding.piece_destinations) against
    plain concatenation: piece k of rank r starts where assemble_pieces puts it."""
    import random
    from jpeg_encoder_b200 import sharding
    rnd = random.Random(7)
    for world, n_scans in ((1, 1), (2, 3), (8, 12), (5, 7)):
        pieces = [[bytes(rnd.getrandbits(8) for _ in range(rnd.randint(0, 40))) for _ in range(n_scans)] for _ in range(world)]
        table = []
        for r in range(world):
            offs = [0]
            for k in range(n_scans):
                offs.append(offs[-1] + len(pieces[r][k]))
            table.append(offs)
        dest, total = sharding.piece_destinations(table)
        want = sharding.assemble_pieces(pieces)
        assert total == len(want)
        for r in range(world):
            for k in range(n_scans):
                assert want[dest[r][k]:dest[r][k] + len(pieces[r][k])] == pieces[r][k]
