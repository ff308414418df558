"""A small FLAC *encoder* for the tests (test infrastructure, never imported by the
package): written independently of the decoder in csrc/flac.cu, from the FLAC format
specification, so that the two check each other.  It can be told which subframe type,
predictor order, Rice partitioning and stereo mode to use for each frame, which is what the
decoder tests need; it makes no attempt to compress well."""
import hashlib
import struct

import numpy as np


class _Bits(object):
    def __init__(self):
        self.acc = 0
        self.n = 0
        self.out = bytearray()

    def put(self, value, width):
        if width == 0:
            return
        value &= (1 << width) - 1
        self.acc = (self.acc << width) | value
        self.n += width
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


def _crc(data, poly, width):
    top = 1 << (width - 1)
    mask = (1 << width) - 1
    c = 0
    for b in data:
        c ^= b << (width - 8)
        for _ in range(8):
            c = ((c << 1) ^ poly) & mask if c & top else (c << 1) & mask
    return c


def _utf8_number(n):
    if n < 0x80:
        return bytes([n])
    cont = []
    payload_bits = 5                      # payload of the lead byte of the 2-byte form
    while True:
        cont.append(0x80 | (n & 0x3F))
        n >>= 6
        if n < (1 << payload_bits):
            break
        payload_bits -= 1
    lead = ((0xFF << (payload_bits + 1)) & 0xFF) | n
    return bytes([lead] + cont[::-1])


def _rice_param(res):
    mean = float(np.mean(np.abs(res))) if len(res) else 0.0
    k = 0
    while (1 << k) < mean + 1 and k < 14:
        k += 1
    return k


def _put_residual(bw, res, blocksize, order, porder, method=0, escape_first=False, rice_k=None):
    bw.put(method, 2)
    bw.put(porder, 4)
    pbits = 4 if method == 0 else 5
    i = 0
    for part in range(1 << porder):
        count = (blocksize >> porder) - (order if part == 0 else 0)
        chunk = [int(v) for v in res[i:i + count]]
        i += count
        if escape_first and part == 0:
            raw = max([1] + [(abs(v) + 1).bit_length() + 1 for v in chunk])
            bw.put((1 << pbits) - 1, pbits)
            bw.put(raw, 5)
            for v in chunk:
                bw.put(v, raw)
            continue
        k = _rice_param(np.asarray(chunk)) if rice_k is None else rice_k
        bw.put(k, pbits)
        for v in chunk:
            u = (v << 1) if v >= 0 else ((-v) << 1) - 1
            bw.unary(u >> k)
            bw.put(u & ((1 << k) - 1), k)
    assert i == len(res)


_FIXED = {0: [], 1: [1], 2: [2, -1], 3: [3, -3, 1], 4: [4, -6, 4, -1]}


def _put_subframe(bw, x, bps, spec):
    """x: python ints of one channel of one block; spec: dict(kind=..., ...)."""
    n = len(x)
    kind = spec.get("kind", "fixed")
    wasted = spec.get("wasted", 0)
    if wasted:
        assert all(v % (1 << wasted) == 0 for v in x)
        x = [v >> wasted for v in x]
    bw.put(0, 1)
    if kind == "constant":
        code, order = 0, 0
    elif kind == "verbatim":
        code, order = 1, 0
    elif kind == "fixed":
        order = spec.get("order", 2)
        code = 8 + order
    elif kind == "lpc":
        coefs = spec["coefs"]
        order = len(coefs)
        code = 31 + order
    else:
        raise ValueError(kind)
    bw.put(code, 6)
    if wasted:
        bw.put(1, 1)
        bw.unary(wasted - 1)
    else:
        bw.put(0, 1)
    b = bps - wasted
    if kind == "constant":
        assert all(v == x[0] for v in x)
        bw.put(x[0], b)
        return
    if kind == "verbatim":
        for v in x:
            bw.put(v, b)
        return
    for v in x[:order]:
        bw.put(v, b)
    if kind == "fixed":
        c = _FIXED[order]
        res = [x[i] - sum(c[j] * x[i - 1 - j] for j in range(order)) for i in range(order, n)]
    else:
        precision, shift = spec.get("precision", 12), spec.get("shift", 9)
        bw.put(precision - 1, 4)
        bw.put(shift, 5)
        for cf in coefs:
            bw.put(cf, precision)
        res = [x[i] - (sum(coefs[j] * x[i - 1 - j] for j in range(order)) >> shift) for i in range(order, n)]
    porder = spec.get("porder", 0)
    if n % (1 << porder) or (n >> porder) <= order:
        porder = 0                                   # a short last block cannot be split evenly
    _put_residual(bw, res, n, order, porder, spec.get("method", 0), spec.get("escape_first", False), spec.get("rice_k"))


def _blocksize_code(n):
    if n == 192:
        return 1, None
    for c in range(2, 6):
        if n == 576 << (c - 2):
            return c, None
    for c in range(8, 16):
        if n == 256 << (c - 8):
            return c, None
    if n <= 256:
        return 6, (n - 1, 8)
    return 7, (n - 1, 16)


def encode(samples, sample_rate, bps=16, blocksize=1152, plan=None, write_md5=True, id3=False):
    """samples: int array [frames, channels].  plan(frame_index, channels) -> dict(stereo=None|'ls'|'sr'|'ms',
    sub=[spec per channel]).  Returns the bytes of a .flac file."""
    samples = np.asarray(samples)
    if samples.ndim == 1:
        samples = samples[:, None]
    frames, nch = samples.shape
    out = bytearray()
    if id3:
        out += b"ID3\x03\x00\x00" + bytes([0, 0, 0, 10]) + b"\x00" * 10
    out += b"fLaC"
    width = (bps + 7) // 8
    raw = samples.astype("<i%d" % (4 if width == 3 else width))
    if width == 3:
        raw = np.frombuffer(raw.tobytes(), np.uint8).reshape(-1, 4)[:, :3]
    md5 = hashlib.md5(raw.tobytes()).digest() if write_md5 else b"\x00" * 16
    info = _Bits()
    info.put(blocksize, 16)
    info.put(blocksize, 16)
    info.put(0, 24)
    info.put(0, 24)
    info.put(sample_rate, 20)
    info.put(nch - 1, 3)
    info.put(bps - 1, 5)
    info.put(frames, 36)
    body = info.bytes() + md5
    out += bytes([0x00]) + struct.pack(">I", len(body))[1:] + body
    pad = b"\x00" * 8
    out += bytes([0x81]) + struct.pack(">I", len(pad))[1:] + pad           # a PADDING block, flagged last
    idx = 0
    for start in range(0, frames, blocksize):
        block = samples[start:start + blocksize].astype(np.int64)
        n = block.shape[0]
        spec = plan(idx, nch) if plan else {"stereo": None, "sub": [{"kind": "fixed", "order": 2}] * nch}
        stereo = spec.get("stereo") if nch == 2 else None
        chans = [[int(v) for v in block[:, c]] for c in range(nch)]
        widths = [bps] * nch
        if stereo == "ls":
            chans[1] = [a - b for a, b in zip(chans[0], chans[1])]
            widths[1] += 1
            ch_code = 8
        elif stereo == "sr":
            chans[0] = [a - b for a, b in zip(chans[0], chans[1])]
            widths[0] += 1
            ch_code = 9
        elif stereo == "ms":
            mid = [(a + b) >> 1 for a, b in zip(chans[0], chans[1])]
            side = [a - b for a, b in zip(chans[0], chans[1])]
            chans = [mid, side]
            widths[1] += 1
            ch_code = 10
        else:
            ch_code = nch - 1
        bw = _Bits()
        bw.put(0x3FFE, 14)
        bw.put(0, 1)
        bw.put(0, 1)                                   # fixed block size stream: the coded number is the frame index
        bcode, bextra = _blocksize_code(n)
        bw.put(bcode, 4)
        bw.put(0, 4)                                   # sample rate: see STREAMINFO
        bw.put(ch_code, 4)
        bw.put({8: 1, 12: 2, 16: 4, 20: 5, 24: 6}.get(bps, 0), 3)
        bw.put(0, 1)
        for byte in _utf8_number(idx):
            bw.put(byte, 8)
        if bextra:
            bw.put(*bextra)
        hdr = bw.bytes()
        bw.put(_crc(hdr, 0x07, 8), 8)
        for c in range(nch):
            _put_subframe(bw, chans[c], widths[c], spec["sub"][c])
        bw.align()
        frame = bw.bytes()
        out += frame + struct.pack(">H", _crc(frame, 0x8005, 16))
        idx += 1
    return bytes(out)
