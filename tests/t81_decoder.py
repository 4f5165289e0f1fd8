"""An independent ITU-T T.81 entropy decoder -- TEST INFRASTRUCTURE ONLY.

Written from the JPEG standard (Annex B syntax, Annex F sequential and Annex G spectral-selection
progressive decoding, Annex C code assignment), NOT from the reference encoder and NOT from the
oracle: it shares no code with either. It stops at the quantized coefficients, so a file can be
checked *exactly* (no IDCT tolerance): the coefficients a conformant decoder recovers from the
file must equal the ones the unit-pinned colour / fDCT / quantizer functions produce.

Also checks what the syntax fixes: marker order, segment lengths, that every restart interval and
scan ends in 1-padding inside its last byte (F.1.2.3), RSTn numbering modulo 8, no stray bytes.
"""
import numpy as np


class JpegSyntaxError(Exception):
    pass


class _Huff:
    """Annex C: canonical codes from BITS / HUFFVAL; decode through a 16-bit look-ahead table."""

    def __init__(self, counts, values):
        self.counts, self.values = list(counts), list(values)
        lut_len = np.zeros(65536, np.uint8)
        lut_sym = np.zeros(65536, np.uint8)
        code, k = 0, 0
        for length in range(1, 17):
            for _ in range(counts[length - 1]):
                lo = code << (16 - length)
                hi = lo + (1 << (16 - length))
                if hi > 65536:
                    raise JpegSyntaxError("Huffman code space overflow")
                lut_len[lo:hi] = length
                lut_sym[lo:hi] = values[k]
                code += 1
                k += 1
            code <<= 1
        self.lut_len = lut_len.tolist()
        self.lut_sym = lut_sym.tolist()


class _Bits:
    """MSB-first reader over one restart interval (already un-stuffed)."""

    def __init__(self, data):
        self.n = len(data)
        self.d = bytes(data) + b"\xff\xff\xff\xff"  # look-ahead padding; consumption is checked at the end
        self.pos = 0  # in bits

    def peek16(self):
        i, o = self.pos >> 3, self.pos & 7
        d = self.d
        return (((d[i] << 16) | (d[i + 1] << 8) | d[i + 2]) >> (8 - o)) & 0xFFFF

    def take(self, n):
        if n == 0:
            return 0
        i, o = self.pos >> 3, self.pos & 7
        d = self.d
        v = ((d[i] << 24) | (d[i + 1] << 16) | (d[i + 2] << 8) | d[i + 3]) >> (32 - o - n)
        self.pos += n
        return v & ((1 << n) - 1)

    def symbol(self, h):
        p = self.peek16()
        length = h.lut_len[p]
        if length == 0:
            raise JpegSyntaxError("invalid Huffman code at bit %d" % self.pos)
        self.pos += length
        return h.lut_sym[p]

    def finish(self):
        """F.1.2.3: the interval ends inside its last byte and the unused bits are all ones."""
        if self.pos > self.n * 8:
            raise JpegSyntaxError("entropy data over-read by %d bits" % (self.pos - self.n * 8))
        if self.n * 8 - self.pos >= 8:
            raise JpegSyntaxError("%d unused bytes at the end of an interval" % ((self.n * 8 - self.pos) // 8))
        rest = self.n * 8 - self.pos
        if rest and self.take(rest) != (1 << rest) - 1:
            raise JpegSyntaxError("padding bits are not ones")


def _extend(v, t):  # F.2.2.1 EXTEND
    return v if t == 0 or v >= (1 << (t - 1)) else v - (1 << t) + 1


class Decoded:
    """What the file says: frame parameters, tables as written, segment list, coefficients.

    coef[c] is (blocks_h, blocks_w, 64) int32 over the MCU-padded grid in zig-zag order; `seen[c]`
    counts how many scans touched each (block, coefficient)."""


def decode(jpg, headers_only=False):
    """headers_only: stop at the first SOS (tables and frame parameters only)."""
    jpg = bytes(jpg)
    r = Decoded()
    r.segments, r.apps, r.qt, r.scans = [], [], {}, []
    r.progressive, r.restart_interval = False, 0
    dc_tabs, ac_tabs = {}, {}
    r.dht = []
    if jpg[:2] != b"\xff\xd8":
        raise JpegSyntaxError("no SOI")
    pos, frame = 2, None
    r.segments.append("SOI")
    while True:
        if pos + 2 > len(jpg) or jpg[pos] != 0xFF:
            raise JpegSyntaxError("marker expected at %d" % pos)
        m = jpg[pos + 1]
        pos += 2
        if m == 0xD9:
            r.segments.append("EOI")
            if pos != len(jpg):
                raise JpegSyntaxError("%d bytes after EOI" % (len(jpg) - pos))
            break
        ln = (jpg[pos] << 8) | jpg[pos + 1]
        body = jpg[pos + 2:pos + ln]
        if len(body) != ln - 2:
            raise JpegSyntaxError("truncated segment %02X" % m)
        pos += ln
        if 0xE0 <= m <= 0xEF:
            r.segments.append("APP%d" % (m - 0xE0))
            r.apps.append((m - 0xE0, body))
        elif m == 0xDB:
            r.segments.append("DQT")
            i = 0
            while i < len(body):
                pq, tq = body[i] >> 4, body[i] & 15
                i += 1
                if pq == 0:
                    vals = list(body[i:i + 64])
                    i += 64
                else:
                    vals = [(body[i + 2 * k] << 8) | body[i + 2 * k + 1] for k in range(64)]
                    i += 128
                r.qt[tq] = (pq, vals)  # zig-zag order, as written
        elif m in (0xC0, 0xC1, 0xC2):
            r.segments.append("SOF%d" % (m - 0xC0))
            if frame is not None:
                raise JpegSyntaxError("second frame header")
            r.progressive = m == 0xC2
            r.precision = body[0]
            r.height, r.width = (body[1] << 8) | body[2], (body[3] << 8) | body[4]
            nf = body[5]
            if len(body) != 6 + 3 * nf:
                raise JpegSyntaxError("SOF length")
            frame = [(body[6 + 3 * k], body[7 + 3 * k] >> 4, body[7 + 3 * k] & 15, body[8 + 3 * k]) for k in range(nf)]
            r.components = frame
            hmax, vmax = max(f[1] for f in frame), max(f[2] for f in frame)
            r.hmax, r.vmax = hmax, vmax
            r.mcu_cols = -(-r.width // (8 * hmax))
            r.mcu_rows = -(-r.height // (8 * vmax))
            r.coef = [np.zeros((r.mcu_rows * f[2], r.mcu_cols * f[1], 64), np.int32) for f in frame]
            r.seen = [np.zeros((r.mcu_rows * f[2], r.mcu_cols * f[1], 64), np.uint8) for f in frame]
        elif m == 0xC4:
            r.segments.append("DHT")
            i = 0
            while i < len(body):
                tc, th = body[i] >> 4, body[i] & 15
                counts = list(body[i + 1:i + 17])
                n = sum(counts)
                values = list(body[i + 17:i + 17 + n])
                if len(values) != n:
                    raise JpegSyntaxError("DHT length")
                i += 17 + n
                (dc_tabs if tc == 0 else ac_tabs)[th] = _Huff(counts, values)
                r.dht.append((tc, th, counts, values))
        elif m == 0xDD:
            r.segments.append("DRI")
            if ln != 4:
                raise JpegSyntaxError("DRI length")
            r.restart_interval = (body[0] << 8) | body[1]
        elif m == 0xDA:
            r.segments.append("SOS")
            if frame is None:
                raise JpegSyntaxError("SOS before SOF")
            if headers_only:
                return r
            ns = body[0]
            sel = [(body[1 + 2 * k], body[2 + 2 * k] >> 4, body[2 + 2 * k] & 15) for k in range(ns)]
            ss, se, ahal = body[1 + 2 * ns], body[2 + 2 * ns], body[3 + 2 * ns]
            if len(body) != 4 + 2 * ns:
                raise JpegSyntaxError("SOS length")
            # entropy-coded data: up to the next marker that is neither RSTn nor a stuffed zero
            intervals, cur, rst_seen = [], bytearray(), []
            while True:
                j = jpg.find(b"\xff", pos)
                if j < 0:
                    raise JpegSyntaxError("entropy data runs off the file")
                cur += jpg[pos:j]
                nxt = jpg[j + 1]
                if nxt == 0x00:
                    cur.append(0xFF)
                    pos = j + 2
                elif 0xD0 <= nxt <= 0xD7:
                    intervals.append(bytes(cur))
                    cur = bytearray()
                    rst_seen.append(nxt - 0xD0)
                    pos = j + 2
                else:
                    intervals.append(bytes(cur))
                    pos = j
                    break
            _decode_scan(r, sel, ss, se, ahal, intervals, rst_seen, dc_tabs, ac_tabs)
            r.scans.append(dict(components=[s[0] for s in sel], ss=ss, se=se, ah=ahal >> 4, al=ahal & 15,
                                tables=[(s[1], s[2]) for s in sel], n_intervals=len(intervals),
                                data_bytes=sum(len(i) for i in intervals)))
        else:
            raise JpegSyntaxError("unexpected marker %02X" % m)
    return r


def _decode_scan(r, sel, ss, se, ahal, intervals, rst_seen, dc_tabs, ac_tabs):
    if ahal != 0:
        raise JpegSyntaxError("successive approximation is not expected here (Ah/Al = %02X)" % ahal)
    if not r.progressive and (ss, se) != (0, 63):
        raise JpegSyntaxError("sequential scan must have Ss=0, Se=63")
    if r.progressive and ss == 0 and se != 0:
        raise JpegSyntaxError("progressive DC scan must have Se=0")
    if r.progressive and ss > 0 and len(sel) != 1:
        raise JpegSyntaxError("progressive AC scans are single-component (G.1.1.1.1)")
    idx = {f[0]: i for i, f in enumerate(r.components)}
    comps = [idx[s[0]] for s in sel]
    for k, n in enumerate(rst_seen):  # RSTm, m counts modulo 8 from 0 (E.1.4 / B.2.1)
        if n != k % 8:
            raise JpegSyntaxError("RST%d where RST%d was due" % (n, k % 8))
    # the units (MCUs) of the scan in order, each a list of (component, block_y, block_x, dc_tab, ac_tab)
    if len(comps) == 1:
        c = comps[0]
        _, h, v, _ = r.components[c]
        bw = -(-(-(-r.width * h // r.hmax)) // 8)   # A.2.3: ceil(ceil(X * Hi / Hmax) / 8)
        bh = -(-(-(-r.height * v // r.vmax)) // 8)
        n_units = bw * bh

        def unit(u):
            return [(c, u // bw, u % bw, sel[0][1], sel[0][2])]
    else:
        n_units = r.mcu_cols * r.mcu_rows

        def unit(u):
            my, mx = divmod(u, r.mcu_cols)
            out = []
            for (cid, td, ta), c in zip(sel, comps):
                _, h, v, _ = r.components[c]
                for yy in range(v):
                    for xx in range(h):
                        out.append((c, my * v + yy, mx * h + xx, td, ta))
            return out
    ri = r.restart_interval
    want = 1 if ri == 0 else -(-n_units // ri)
    if len(intervals) != want:
        raise JpegSyntaxError("%d restart intervals in the scan, %d expected" % (len(intervals), want))
    u = 0
    for data in intervals:
        bits = _Bits(data)
        pred = {c: 0 for c in comps}
        eobrun = 0
        last = n_units if ri == 0 else min(n_units, u + ri)
        while u < last:
            for (c, by, bx, td, ta) in unit(u):
                blk = r.coef[c][by, bx]
                seen = r.seen[c][by, bx]
                k = ss
                if ss == 0:
                    t = bits.symbol(dc_tabs[td])
                    if t > 11:
                        raise JpegSyntaxError("DC category %d" % t)
                    pred[c] += _extend(bits.take(t), t)
                    blk[0] = pred[c]
                    seen[0] += 1
                    k = 1
                if se > 0:
                    seen[k:se + 1] += 1
                    if eobrun > 0:
                        eobrun -= 1
                        continue
                    h = ac_tabs[ta]
                    while k <= se:
                        rs = bits.symbol(h)
                        run, size = rs >> 4, rs & 15
                        if size == 0:
                            if run == 15:
                                k += 16
                                continue
                            if run == 0:
                                break
                            if not r.progressive:
                                raise JpegSyntaxError("EOBn in a sequential scan")
                            eobrun = (1 << run) + bits.take(run) - 1
                            break
                        k += run
                        if k > se:
                            raise JpegSyntaxError("run past the end of the band")
                        blk[k] = _extend(bits.take(size), size)
                        k += 1
            u += 1
        if eobrun:
            raise JpegSyntaxError("EOB run crosses a restart interval")
        bits.finish()
    if u != n_units:
        raise JpegSyntaxError("scan ended after %d of %d units" % (u, n_units))
