"""How the encode path is spread over the GPUs of one box.

* Batches shard by image: rank r owns images [lo, hi) -- no data-path collective at all.
* One very large image shards by strips of whole MCU rows whose boundaries are restart boundaries
  of every scan. Each rank encodes its strip; the only exchange is the gather of the strips'
  per-scan byte pieces to rank 0 (sizes first, then one grouped batch of variable-length sends over
  NCCL/NVLink), where they are concatenated scan-major. With optimized Huffman tables the strips
  first all-reduce their symbol histograms (the tables describe the whole image).

The exchange functions work on whatever device the tensors live on: NCCL with CUDA tensors on the
B200 box, gloo with CPU tensors in the world_size-2 CPU tests.
"""
import torch
import torch.distributed as dist


def shard_batch(n_images, world, rank):
    """Contiguous, balanced image range of `rank`: sizes differ by at most one."""
    base, extra = divmod(n_images, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def assemble_pieces(pieces_by_strip):
    """pieces_by_strip[s][k] = bytes of scan k produced by strip s. The file is scan-major:
    for each scan, strips in order (strip 0's piece carries the header / SOS, the last strip's last
    piece the EOI)."""
    n_scans = len(pieces_by_strip[0])
    out = bytearray()
    for k in range(n_scans):
        for strip in pieces_by_strip:
            out += strip[k]
    return bytes(out)


def split_pieces(buf, offsets):
    return [bytes(buf[offsets[k]:offsets[k + 1]]) for k in range(len(offsets) - 1)]


def exchange_strip_histograms(hist, edge_dc, device, group=None):
    """Optimized Huffman tables with strips (include/jpegenc_b200.h, steps 2): the strips' symbol histograms
    are summed with one all-reduce (2 x 2 x 257 u32; NCCL over NVLink on the GPU box) and the DC values at
    the strip edges are all-gathered in rank order. Returns (hist_sum list, edge_dc list of world * 8)."""
    h = torch.tensor(hist, dtype=torch.int64, device=device)  # NCCL has no u32 sum; counts fit easily
    dist.all_reduce(h, op=dist.ReduceOp.SUM, group=group)
    e = torch.tensor(edge_dc, dtype=torch.int32, device=device)
    world = dist.get_world_size(group)
    table = torch.empty(world * 8, dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(table, e, group=group)
    return h.cpu().tolist(), table.cpu().tolist()


def gather_strip_pieces(local_bytes, piece_offsets, rank, world, device, group=None):
    """local_bytes: uint8 tensor on `device` holding this rank's pieces back to back.
    Returns on rank 0 the assembled uint8 tensor of the whole file (scan-major), on other ranks None.
    Communication: one all_gather of the (n_scans + 1) offsets, then one batch of point-to-point
    transfers into rank 0 (a single grouped NCCL send/recv over NVLink; exact sizes, no padding)."""
    n = len(piece_offsets)
    mine = torch.tensor(piece_offsets, dtype=torch.int64, device=device)
    table = torch.empty(world * n, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(table, mine, group=group)
    table = table.cpu().view(world, n).tolist()
    if rank == 0:
        bufs = [local_bytes] + [torch.empty(table[r][-1], dtype=torch.uint8, device=device) for r in range(1, world)]
        ops = [dist.P2POp(dist.irecv, bufs[r], r, group) for r in range(1, world) if table[r][-1] > 0]
        if ops:
            for q in dist.batch_isend_irecv(ops):
                q.wait()
        parts = [bufs[r][table[r][k]:table[r][k + 1]] for k in range(n - 1) for r in range(world)]
        return torch.cat(parts)
    if piece_offsets[-1] > 0:
        for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, local_bytes[:piece_offsets[-1]], 0, group)]):
            q.wait()
    return None
