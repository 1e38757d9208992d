"""TEST INFRASTRUCTURE -- CPU restatement of the JPEG frame decode under the reference's readers.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu legs may import this; nothing under spatialaudiogen_b200/ does.

What it restates.  The reference reads every video / flow frame with `scipy.misc.imread` (feeder.py:120-127), i.e. PIL on
top of the un-vendored third-party library libjpeg (libjpeg-turbo in this image: PIL.features.version('jpg') = 6.2 API,
turbo).  The arithmetic therefore lives in libjpeg's baseline decoder with its default settings -- JDCT_ISLOW, fancy
upsampling -- whose published algorithms are restated here:

    entropy decoding      jdhuff.c   decode_mcu (canonical Huffman codes of the DHT segments, HUFF_EXTEND, zig-zag order)
    dequantise + IDCT     jidctint.c jpeg_idct_islow (CONST_BITS 13, PASS1_BITS 2, the 12 FIX_* constants)
    chroma upsampling     jdsample.c h2v2_fancy_upsample / h2v1_fancy_upsample (triangle filter), edge rows replicated as
                          jdmainct.c does
    colour conversion     jdcolor.c  build_ycc_rgb_table / ycc_rgb_convert (16-bit fixed point)

Pinned by tests/test_jpeg.py against PIL's own decode (bit-exact) of JPEG files written by PIL at several qualities /
subsamplings / sizes, and of the reference's own photographs (pyutils/tflib/models/image/test_images/*.jp*g, staged under
tests/golden/_ref/ by build()).

Scope: baseline sequential (SOF0), 8 bit, 1 or 3 components in one interleaved scan, sampling 4:4:4 / 4:2:2 / 4:2:0,
restart intervals.  Progressive files raise ValueError (the reference's frames are written by skimage.io.imsave, i.e. PIL's
encoder with its defaults -- baseline, 4:2:0, quality 75 -- scraping/preprocess.py:141-143 and :198: the same encoder the tests
write their files with)."""
import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                   35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55,
                   62, 63])


def parse(data):
    """Marker segments of a baseline file -> dict(width, height, comps=[(id, h, v, tq)], qt={id: (64,) natural order},
    huff={(cls, id): (counts[16], symbols)}, restart_interval, scan=[(comp index, dc table, ac table)], ecs=bytes)."""
    d = bytes(data)
    if d[:2] != b'\xff\xd8':
        raise ValueError('not a JPEG file (no SOI)')
    out = {'qt': {}, 'huff': {}, 'restart_interval': 0}
    jfif, adobe = False, None
    p = 2
    while True:
        while d[p] != 0xFF:
            p += 1
        while d[p] == 0xFF:
            p += 1
        m = d[p]
        p += 1
        if m == 0xD9:
            raise ValueError('EOI before SOS')
        if m in (0x01,) or 0xD0 <= m <= 0xD7:
            continue
        n = (d[p] << 8) | d[p + 1]
        seg = d[p + 2:p + n]
        p += n
        if m == 0xDB:
            q = 0
            while q < len(seg):
                pq, tq = seg[q] >> 4, seg[q] & 15
                if pq == 0:
                    vals = np.frombuffer(seg[q + 1:q + 65], np.uint8).astype(np.int32)
                    q += 65
                else:
                    vals = np.frombuffer(seg[q + 1:q + 129], '>u2').astype(np.int32)
                    q += 129
                t = np.zeros(64, np.int32)
                t[ZIGZAG] = vals
                out['qt'][tq] = t
        elif m == 0xC0 or m == 0xC1:
            if seg[0] != 8:
                raise ValueError('only 8-bit samples')
            out['height'], out['width'] = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4]
            out['comps'] = [(seg[6 + 3 * i], seg[7 + 3 * i] >> 4, seg[7 + 3 * i] & 15, seg[8 + 3 * i]) for i in range(seg[5])]
        elif m in (0xC2, 0xC3, 0xC5, 0xC6, 0xC7, 0xC9, 0xCA, 0xCB, 0xCD, 0xCE, 0xCF):
            raise ValueError('only baseline sequential JPEG (SOF0) is supported, found SOF%d' % (m - 0xC0))
        elif m == 0xC4:
            q = 0
            while q < len(seg):
                cls, tid = seg[q] >> 4, seg[q] & 15
                counts = list(seg[q + 1:q + 17])
                ns = sum(counts)
                out['huff'][(cls, tid)] = (counts, list(seg[q + 17:q + 17 + ns]))
                q += 17 + ns
        elif m == 0xE0 and seg[:5] == b'JFIF\0':
            jfif = True
        elif m == 0xEE and seg[:5] == b'Adobe' and len(seg) >= 12:
            adobe = seg[11]
        elif m == 0xDD:
            out['restart_interval'] = (seg[0] << 8) | seg[1]
        elif m == 0xDA:
            ns = seg[0]
            ids = [c[0] for c in out['comps']]
            out['scan'] = [(ids.index(seg[1 + 2 * i]), seg[2 + 2 * i] >> 4, seg[2 + 2 * i] & 15) for i in range(ns)]
            if ns != len(out['comps']):
                raise ValueError('only one interleaved scan with all components is supported')
            # jdapimin.c default_decompress_parms: without a JFIF marker, Adobe transform 0 or component ids R, G, B mean RGB samples
            if len(ids) == 3 and not jfif and (adobe == 0 or (adobe is None and ids == [82, 71, 66])):
                raise ValueError('the file stores RGB samples: only YCbCr and grey files are supported')
            out['ecs'] = d[p:]
            return out


def _decode_table(counts, symbols):
    """jdhuff.c jpeg_make_d_derived_tbl: {(length, code): symbol} of the canonical code."""
    table, code, k = {}, 0, 0
    for length in range(1, 17):
        for _ in range(counts[length - 1]):
            table[(length, code)] = symbols[k]
            code += 1
            k += 1
        code <<= 1
    return table


class _Bits(object):
    def __init__(self, ecs):
        self.d, self.p, self.acc, self.n = ecs, 0, 0, 0

    def _fill(self):
        d = self.d
        if self.p < len(d):
            b = d[self.p]
            if b == 0xFF:
                nxt = d[self.p + 1] if self.p + 1 < len(d) else 0xD9
                if nxt == 0:
                    self.p += 2
                else:                      # a marker: the entropy-coded segment ended; feed zeros (jdhuff.c: "insert_fake_zeros")
                    b = 0
            else:
                self.p += 1
        else:
            b = 0
        self.acc = ((self.acc << 8) | b) & 0xFFFFFFFF
        self.n += 8

    def get(self, k):
        while self.n < k:
            self._fill()
        self.n -= k
        return (self.acc >> self.n) & ((1 << k) - 1)

    def restart(self):
        self.acc, self.n = 0, 0            # discard the partial byte, then skip the RSTn marker
        d = self.d
        while self.p + 1 < len(d) and not (d[self.p] == 0xFF and 0xD0 <= d[self.p + 1] <= 0xD7):
            self.p += 1
        self.p += 2


def _huff(bits, table):
    code = 0
    for length in range(1, 17):
        code = (code << 1) | bits.get(1)
        s = table.get((length, code))
        if s is not None:
            return s
    raise ValueError('bad Huffman code')


def _extend(r, s):
    return r if r >= (1 << (s - 1)) else r - (1 << s) + 1        # jdhuff.h HUFF_EXTEND


def coefficients(hdr):
    """Quantised coefficients per component: list of (blocks_high, blocks_wide, 64) int16 arrays in natural order (the block
    grid is padded to whole MCUs like libjpeg's coefficient buffers)."""
    comps = hdr['comps']
    hmax, vmax = max(c[1] for c in comps), max(c[2] for c in comps)
    mcux, mcuy = -(-hdr['width'] // (8 * hmax)), -(-hdr['height'] // (8 * vmax))
    if len(comps) == 1:                    # a single-component scan is never interleaved: MCU = one block
        hmax = vmax = 1
        comps = [(comps[0][0], 1, 1, comps[0][3])]
        mcux, mcuy = -(-hdr['width'] // 8), -(-hdr['height'] // 8)
    coef = [np.zeros((mcuy * c[2], mcux * c[1], 64), np.int16) for c in comps]
    tabs = {k: _decode_table(*v) for k, v in hdr['huff'].items()}
    bits = _Bits(hdr['ecs'])
    pred = [0] * len(comps)
    ri, left = hdr['restart_interval'], hdr['restart_interval']
    for my in range(mcuy):
        for mx in range(mcux):
            if ri and left == 0:
                bits.restart()
                pred = [0] * len(comps)
                left = ri
            for ci, td, ta in hdr['scan']:
                _, h, v, _ = comps[ci]
                for by in range(v):
                    for bx in range(h):
                        blk = coef[ci][my * v + by, mx * h + bx]
                        s = _huff(bits, tabs[(0, td)])
                        if s:
                            pred[ci] += _extend(bits.get(s), s)
                        blk[0] = pred[ci]
                        k = 1
                        while k < 64:
                            rs = _huff(bits, tabs[(1, ta)])
                            r, s = rs >> 4, rs & 15
                            if s:
                                k += r
                                blk[ZIGZAG[k & 63]] = _extend(bits.get(s), s)
                            elif r == 15:
                                k += 15
                            else:
                                break
                            k += 1
            left -= 1
    return coef


# ---- jidctint.c -----------------------------------------------------------------------------------------------------
_C = dict(f0298=2446, f0390=3196, f0541=4433, f0765=6270, f0899=7373, f1175=9633, f1501=12299, f1847=15137, f1961=16069, f2053=16819,
          f2562=20995, f3072=25172)


def _idct_1d(x, shift):
    """One pass of jpeg_idct_islow over the LAST axis of x (..., 8) int64; DESCALE by `shift`."""
    c = _C
    z2, z3 = x[..., 2], x[..., 6]
    z1 = (z2 + z3) * c['f0541']
    tmp2 = z1 - z3 * c['f1847']
    tmp3 = z1 + z2 * c['f0765']
    z2, z3 = x[..., 0], x[..., 4]
    tmp0, tmp1 = (z2 + z3) << 13, (z2 - z3) << 13
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    tmp0, tmp1, tmp2, tmp3 = x[..., 7], x[..., 5], x[..., 3], x[..., 1]
    z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
    z5 = (z3 + z4) * c['f1175']
    tmp0, tmp1, tmp2, tmp3 = tmp0 * c['f0298'], tmp1 * c['f2053'], tmp2 * c['f3072'], tmp3 * c['f1501']
    z1, z2, z3, z4 = -z1 * c['f0899'], -z2 * c['f2562'], -z3 * c['f1961'] + z5, -z4 * c['f0390'] + z5
    tmp0, tmp1, tmp2, tmp3 = tmp0 + z1 + z3, tmp1 + z2 + z4, tmp2 + z2 + z3, tmp3 + z1 + z4
    out = np.stack([tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0, tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2, tmp10 - tmp3], -1)
    return (out + (1 << (shift - 1))) >> shift


def idct_blocks(coef, qt):
    """(by, bx, 64) int16 quantised coefficients x (64,) quantiser -> (by*8, bx*8) uint8 samples."""
    by, bx = coef.shape[:2]
    x = (coef.astype(np.int64) * qt.astype(np.int64)).reshape(by, bx, 8, 8)         # [row][col]
    ws = _idct_1d(x.transpose(0, 1, 3, 2), 13 - 2).transpose(0, 1, 3, 2)           # pass 1: down the columns
    y = _idct_1d(ws, 13 + 2 + 3)                                                    # pass 2: along the rows
    y = (((y & 1023) ^ 512) - 512) + 128                                            # range_limit[x & RANGE_MASK] (10-bit wrap)
    return np.clip(y, 0, 255).astype(np.uint8).transpose(0, 2, 1, 3).reshape(by * 8, bx * 8)


# ---- jdsample.c -------------------------------------------------------------------------------------------------------
def _h2_fancy(row_sums, bias_even, bias_odd, shift, edge_mul):
    """Horizontal triangle filter on (rows, w) sums -> (rows, 2w)."""
    s = row_sums.astype(np.int64)
    w = s.shape[1]
    left = np.concatenate((s[:, :1], s[:, :-1]), 1)
    right = np.concatenate((s[:, 1:], s[:, -1:]), 1)
    even = (s * 3 + left + bias_even) >> shift
    odd = (s * 3 + right + bias_odd) >> shift
    even[:, 0] = (s[:, 0] * edge_mul + bias_even) >> shift if edge_mul else s[:, 0]
    odd[:, -1] = (s[:, -1] * edge_mul + bias_odd) >> shift if edge_mul else s[:, -1]
    out = np.empty((s.shape[0], 2 * w), np.int64)
    out[:, 0::2], out[:, 1::2] = even, odd
    return out


def upsample(plane, h, v, hmax, vmax, width, height):
    """Component samples (padded block grid) -> (height, width) at full resolution, the way libjpeg's default (fancy) upsampler
    does for this sampling ratio."""
    dw, dh = -(-width * h // hmax), -(-height * v // vmax)      # downsampled_width / _height: the real samples
    p = plane[:dh, :dw].astype(np.int64)
    if h == hmax and v == vmax:
        return p[:height, :width].astype(np.uint8)
    if hmax == 2 * h and dw <= 2:                                # jinit_upsampler: fancy only if downsampled_width > 2, else replication
        rep = np.repeat(p, 2, axis=1)
        if vmax == 2 * v:
            rep = np.repeat(rep, 2, axis=0)
        elif v != vmax:
            raise ValueError('sampling %dx%d of %dx%d is not supported' % (h, v, hmax, vmax))
        return rep[:height, :width].astype(np.uint8)
    if hmax == 2 * h and v == vmax:                              # h2v1_fancy_upsample
        out = _h2_fancy(p, 1, 2, 2, 0)
        return out[:height, :width].astype(np.uint8)
    if hmax == 2 * h and vmax == 2 * v:                          # h2v2_fancy_upsample
        above = np.concatenate((p[:1], p[:-1]), 0)               # jdmainct.c replicates the first / last real row
        below = np.concatenate((p[1:], p[-1:]), 0)
        out = np.empty((2 * dh, 2 * dw), np.int64)
        out[0::2] = _h2_fancy(p * 3 + above, 8, 7, 4, 4)
        out[1::2] = _h2_fancy(p * 3 + below, 8, 7, 4, 4)
        return out[:height, :width].astype(np.uint8)
    raise ValueError('sampling %dx%d of %dx%d is not supported' % (h, v, hmax, vmax))


# ---- jdcolor.c --------------------------------------------------------------------------------------------------------
def ycc_to_rgb(y, cb, cr):
    x = np.arange(256, dtype=np.int64) - 128
    cr_r = (91881 * x + 32768) >> 16
    cb_b = (116130 * x + 32768) >> 16
    cr_g = -46802 * x
    cb_g = -22554 * x + 32768
    y = y.astype(np.int64)
    r = y + cr_r[cr]
    g = y + ((cb_g[cb] + cr_g[cr]) >> 16)
    b = y + cb_b[cb]
    return np.clip(np.stack((r, g, b), -1), 0, 255).astype(np.uint8)


def decode(data):
    """bytes of a baseline JPEG file -> (height, width, 3) uint8 RGB, as `np.asarray(PIL.Image.open(f).convert('RGB'))`."""
    hdr = parse(data)
    coef = coefficients(hdr)
    comps = hdr['comps']
    W, H = hdr['width'], hdr['height']
    if len(comps) == 1:
        g = idct_blocks(coef[0], hdr['qt'][comps[0][3]])[:H, :W]
        return np.stack((g, g, g), -1)
    hmax, vmax = max(c[1] for c in comps), max(c[2] for c in comps)
    planes = [upsample(idct_blocks(coef[i], hdr['qt'][c[3]]), c[1], c[2], hmax, vmax, W, H) for i, c in enumerate(comps)]
    return ycc_to_rgb(*planes)
