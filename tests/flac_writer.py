"""A small FLAC *encoder* used only to make test vectors for csrc/flac.cu (there is no flac / soundfile in the image).

It writes valid streams that exercise the decoder paths the reference's own (mono, libFLAC-made) fixtures do not reach:
every subframe type, both Rice variants with escape partitions, wasted bits, the three stereo decorrelations, odd block
sizes through the 8/16-bit header fields, 8/16/24-bit samples.  Compression quality is irrelevant."""
import hashlib
import struct

import numpy as np


class Bits:
    def __init__(self):
        self.acc = 0
        self.n = 0
        self.out = bytearray()

    def put(self, v, k):
        if k == 0:
            return
        self.acc = (self.acc << k) | (int(v) & ((1 << k) - 1))
        self.n += k
        while self.n >= 8:
            self.n -= 8
            self.out.append((self.acc >> self.n) & 0xFF)
        self.acc &= (1 << self.n) - 1

    def unary(self, q):
        while q >= 32:
            self.put(0, 32)
            q -= 32
        self.put(1, q + 1)

    def align(self):
        if self.n:
            self.put(0, 8 - self.n)

    def bytes(self):
        assert self.n == 0
        return bytes(self.out)


def crc8(d):
    c = 0
    for b in d:
        c ^= b
        for _ in range(8):
            c = ((c << 1) ^ 0x07) & 0xFF if c & 0x80 else (c << 1) & 0xFF
    return c


def crc16(d):
    c = 0
    for b in d:
        c ^= b << 8
        for _ in range(8):
            c = ((c << 1) ^ 0x8005) & 0xFFFF if c & 0x8000 else (c << 1) & 0xFFFF
    return c


def _utf8(n):
    """the UTF-8-like variable-length number of the frame header (up to 36 bits)"""
    if n < 0x80:
        return bytes([n])
    for nbytes, limit in ((2, 1 << 11), (3, 1 << 16), (4, 1 << 21), (5, 1 << 26), (6, 1 << 31), (7, 1 << 36)):
        if n < limit:
            cont = [0x80 | ((n >> (6 * i)) & 0x3F) for i in range(nbytes - 1)][::-1]
            lead = ((0xFF << (8 - nbytes)) & 0xFF) | (n >> (6 * (nbytes - 1)))
            return bytes([lead] + cont)
    raise ValueError("number too large")


def _residual(bw, res, order, blocksize, method, po, escape_parts=()):
    bw.put(method, 2)
    bw.put(po, 4)
    pbits = 4 if method == 0 else 5
    idx = 0
    for part in range(1 << po):
        count = (blocksize >> po) - (order if part == 0 else 0) if po else blocksize - order
        seg = [int(v) for v in res[idx : idx + count]]
        idx += count
        if part in escape_parts:
            width = max([1] + [(abs(v) if v >= 0 else abs(v + 1)).bit_length() + 1 for v in seg])
            bw.put((1 << pbits) - 1, pbits)
            bw.put(width, 5)
            for v in seg:
                bw.put(v, width)
            continue
        zz = [(v << 1) if v >= 0 else ((-v) << 1) - 1 for v in seg]
        mean = (sum(zz) // max(len(zz), 1)) if zz else 0
        k = min(max(mean.bit_length() - 1, 0), (1 << pbits) - 2)
        bw.put(k, pbits)
        for u in zz:
            bw.unary(u >> k)
            bw.put(u & ((1 << k) - 1), k)
    assert idx == len(res)


FIXED = {0: [], 1: [1], 2: [2, -1], 3: [3, -3, 1], 4: [4, -6, 4, -1]}


def _subframe(bw, x, bps, spec):
    """spec: dict(kind='constant'|'verbatim'|'fixed'|'lpc', order=, coefs=, shift=, prec=, method=, po=, escape=, wasted=)"""
    x = [int(v) for v in x]
    n = len(x)
    wasted = spec.get("wasted", 0)
    if wasted:
        assert all(v % (1 << wasted) == 0 for v in x)
        x = [v >> wasted for v in x]
        bps -= wasted
    kind = spec["kind"]
    order = spec.get("order", 0)
    code = {"constant": 0, "verbatim": 1}.get(kind)
    if kind == "fixed":
        code = 8 + order
    if kind == "lpc":
        code = 32 + order - 1
    bw.put(0, 1)
    bw.put(code, 6)
    if wasted:
        bw.put(1, 1)
        bw.unary(wasted - 1)
    else:
        bw.put(0, 1)
    if kind == "constant":
        assert len(set(x)) == 1
        bw.put(x[0], bps)
        return
    if kind == "verbatim":
        for v in x:
            bw.put(v, bps)
        return
    for v in x[:order]:
        bw.put(v, bps)
    if kind == "fixed":
        c = FIXED[order]
        res = [x[i] - sum(c[j] * x[i - 1 - j] for j in range(order)) for i in range(order, n)]
    else:
        c, shift, prec = spec["coefs"], spec["shift"], spec["prec"]
        bw.put(prec - 1, 4)
        bw.put(shift, 5)
        for v in c:
            bw.put(v, prec)
        res = [x[i] - (sum(c[j] * x[i - 1 - j] for j in range(order)) >> shift) for i in range(order, n)]
    _residual(bw, res, order, n, spec.get("method", 0), spec.get("po", 0), spec.get("escape", ()))


def encode(pcm, bps=16, sample_rate=16000, blocks=None, stereo_modes=None, specs=None, md5=True):
    """pcm: int array [n] or [n, ch].  blocks: list of block sizes (last may be short).  stereo_modes: per block one of
    'indep', 'ls', 'sr', 'ms'.  specs: per block, per channel subframe spec (see _subframe)."""
    pcm = np.asarray(pcm, np.int64)
    if pcm.ndim == 1:
        pcm = pcm[:, None]
    n, ch = pcm.shape
    blocks = blocks or [4096] * (n // 4096) + ([n % 4096] if n % 4096 else [])
    assert sum(blocks) == n
    nbytes = (bps + 7) // 8
    raw = pcm.astype("<i8").view(np.uint8).reshape(-1, 8)[:, :nbytes].tobytes()
    info = Bits()
    info.put(min(blocks[:-1] or blocks), 16)
    info.put(max(blocks), 16)
    info.put(0, 24)
    info.put(0, 24)
    info.put(sample_rate, 20)
    info.put(ch - 1, 3)
    info.put(bps - 1, 5)
    info.put(n, 36)
    stream = bytearray(b"fLaC" + bytes([0x80 | 0]) + struct.pack(">I", 34)[1:] + info.bytes() +
                       (hashlib.md5(raw).digest() if md5 else bytes(16)))
    pos = 0
    for bi, bs in enumerate(blocks):
        x = pcm[pos : pos + bs]
        pos += bs
        mode = (stereo_modes or ["indep"] * len(blocks))[bi]
        bw = Bits()
        bw.put(0x3FFE, 14)
        bw.put(0, 1)
        bw.put(0, 1)  # fixed block size stream: frame numbers
        std = {192: 1, 576: 2, 1152: 3, 2304: 4, 4608: 5, 256: 8, 512: 9, 1024: 10, 2048: 11, 4096: 12, 8192: 13,
               16384: 14, 32768: 15}
        bcode = std.get(bs, 6 if bs <= 256 else 7)
        bw.put(bcode, 4)
        bw.put(0, 4)  # sample rate: from STREAMINFO
        bw.put({"indep": ch - 1, "ls": 8, "sr": 9, "ms": 10}[mode], 4)
        bw.put({8: 1, 12: 2, 16: 4, 20: 5, 24: 6}.get(bps, 0), 3)
        bw.put(0, 1)
        for b in _utf8(bi):
            bw.put(b, 8)
        if bcode == 6:
            bw.put(bs - 1, 8)
        elif bcode == 7:
            bw.put(bs - 1, 16)
        bw.put(crc8(bw.bytes()), 8)
        chans = [x[:, c] for c in range(ch)]
        widths = [bps] * ch
        if mode == "ls":
            chans = [x[:, 0], x[:, 0] - x[:, 1]]
            widths = [bps, bps + 1]
        elif mode == "sr":
            chans = [x[:, 0] - x[:, 1], x[:, 1]]
            widths = [bps + 1, bps]
        elif mode == "ms":
            chans = [(x[:, 0] + x[:, 1]) >> 1, x[:, 0] - x[:, 1]]
            widths = [bps, bps + 1]
        for c in range(ch):
            spec = (specs[bi][c] if specs else {"kind": "fixed", "order": 2, "po": 0}) if bs > 4 else {"kind": "verbatim"}
            _subframe(bw, chans[c], widths[c], spec)
        bw.align()
        body = bw.bytes()
        stream += body + struct.pack(">H", crc16(body))
    return bytes(stream)
