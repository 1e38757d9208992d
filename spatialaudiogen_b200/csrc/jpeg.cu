// Baseline JPEG frame decode for the readers: the reference reads every video / flow frame with
// scipy.misc.imread (feeder.py:120-127) = PIL over libjpeg with its default settings (JDCT_ISLOW, fancy upsampling).
//
//   host   : marker parsing + Huffman decoding of the entropy-coded segment (a serial bit stream per file; files of a
//            batch are decoded by a small pool of threads) straight into pinned staging -- quantised coefficients, 2 bytes
//            each, the same number of bytes per 4:2:0 frame as its RGB pixels
//   device : jpeg_idct_kernel        dequantise + 8x8 inverse DCT   (libjpeg jidctint.c jpeg_idct_islow, bit-exact)
//            jpeg_rgb_kernel         chroma upsampling (jdsample.c h2v2 / h2v1 fancy triangle filters, edge rows replicated
//                                    like jdmainct.c) + YCbCr -> RGB (jdcolor.c 16-bit fixed point), written as the uint8
//                                    (n, H, W, 3) frames that sag_forward_frames ingests
// Results are bit-identical to PIL's decode (tests/test_jpeg.py; oracle/jpeg_oracle.py restates the same algorithms).
// Scope: SOF0, 8 bit, 1 or 3 components in one interleaved scan, chroma sampled 1x1 / 2x1 / 2x2, restart intervals.
#include "common.cuh"
#include <atomic>
#include <thread>
#include <algorithm>
#include <functional>
#include <memory>

namespace sag {
namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct HuffTable {            // canonical code of one DHT table (jdhuff.c jpeg_make_d_derived_tbl)
  bool present = false;
  uint8_t symbols[256];
  int32_t maxcode[18];        // largest code of each length (-1: none)
  int32_t valoffset[17];      // symbol index of the first code of each length minus that code
  uint16_t look[512];         // 9-bit lookahead: (length << 8) | symbol, 0 = longer than 9 bits
};

struct Component { int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0; };

struct Header {
  int width = 0, height = 0, ncomp = 0;
  Component comp[3];
  uint16_t qt[4][64];          // natural order
  bool qt_present[4] = {false, false, false, false};
  HuffTable dc[4], ac[4];
  int restart_interval = 0;
  const uint8_t* ecs = nullptr;   // entropy-coded segment
  size_t ecs_size = 0;
  int hmax = 1, vmax = 1, mcux = 0, mcuy = 0;
  int bw[3], bh[3];            // block grid of each component (whole MCUs)
};

int build_table(HuffTable& t, const uint8_t* counts, const uint8_t* symbols, int nsym) {
  t.present = true;
  memcpy(t.symbols, symbols, nsym);
  memset(t.look, 0, sizeof(t.look));
  int code = 0, k = 0;
  for (int len = 1; len <= 16; ++len) {
    t.valoffset[len] = k - code;
    if (code + counts[len - 1] > (1 << len) || k + counts[len - 1] > nsym) return SAG_EINVAL;   // more codes than the length has / symbols than listed
    for (int i = 0; i < counts[len - 1]; ++i, ++code, ++k) {
      if (len <= 9) {
        const int lo = code << (9 - len);
        for (int f = 0; f < (1 << (9 - len)); ++f) t.look[lo + f] = (uint16_t)((len << 8) | symbols[k]);
      }
    }
    t.maxcode[len] = counts[len - 1] ? code - 1 : -1;
    if (code > (1 << len)) return SAG_EINVAL;
    code <<= 1;
  }
  t.maxcode[17] = 0x7fffffff;
  return SAG_OK;
}

int parse_header(const uint8_t* d, size_t n, Header* h) {
  SAG_REQUIRE(n >= 4 && d[0] == 0xFF && d[1] == 0xD8, SAG_EINVAL, "jpeg: not a JPEG file (no SOI marker)");
  size_t p = 2;
  bool have_sof = false, saw_jfif = false;
  int adobe_transform = -1;                  // APP14 "Adobe": 0 = components are RGB / CMYK as they are, 1 = YCbCr, 2 = YCCK
  for (;;) {
    while (p < n && d[p] != 0xFF) ++p;
    while (p < n && d[p] == 0xFF) ++p;
    SAG_REQUIRE(p < n, SAG_EINVAL, "jpeg: file ends before the scan");
    const int m = d[p++];
    if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
    SAG_REQUIRE(m != 0xD9, SAG_EINVAL, "jpeg: EOI before the scan");
    SAG_REQUIRE(p + 2 <= n, SAG_EINVAL, "jpeg: truncated marker segment");
    const size_t len = ((size_t)d[p] << 8) | d[p + 1];
    SAG_REQUIRE(len >= 2 && p + len <= n, SAG_EINVAL, "jpeg: truncated marker segment");
    const uint8_t* s = d + p + 2;
    const size_t sl = len - 2;
    p += len;
    if (m == 0xDB) {
      size_t q = 0;
      while (q < sl) {
        const int pq = s[q] >> 4, tq = s[q] & 15;
        SAG_REQUIRE(tq < 4 && q + 1 + (pq ? 128 : 64) <= sl, SAG_EINVAL, "jpeg: bad DQT segment");
        for (int i = 0; i < 64; ++i)
          h->qt[tq][kZigzag[i]] = pq ? (uint16_t)((s[q + 1 + 2 * i] << 8) | s[q + 2 + 2 * i]) : s[q + 1 + i];
        h->qt_present[tq] = true;
        q += 1 + (pq ? 128 : 64);
      }
    } else if (m == 0xC0 || m == 0xC1) {
      SAG_REQUIRE(sl >= 6 && s[0] == 8, SAG_EUNSUPPORTED, "jpeg: only 8-bit samples are supported");
      h->height = (s[1] << 8) | s[2];
      h->width = (s[3] << 8) | s[4];
      h->ncomp = s[5];
      SAG_REQUIRE((h->ncomp == 1 || h->ncomp == 3) && sl >= (size_t)6 + 3 * h->ncomp, SAG_EUNSUPPORTED,
                  "jpeg: %d components (1 or 3 supported)", h->ncomp);
      for (int i = 0; i < h->ncomp; ++i) {
        h->comp[i].id = s[6 + 3 * i];
        h->comp[i].h = s[7 + 3 * i] >> 4;
        h->comp[i].v = s[7 + 3 * i] & 15;
        h->comp[i].tq = s[8 + 3 * i];
        SAG_REQUIRE(h->comp[i].tq < 4 && h->comp[i].h >= 1 && h->comp[i].v >= 1, SAG_EINVAL, "jpeg: bad SOF segment");
      }
      have_sof = true;
    } else if (m >= 0xC2 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
      SAG_REQUIRE(false, SAG_EUNSUPPORTED, "jpeg: only baseline sequential files (SOF0) are supported, found SOF%d", m - 0xC0);
    } else if (m == 0xC4) {
      size_t q = 0;
      while (q < sl) {
        SAG_REQUIRE(q + 17 <= sl, SAG_EINVAL, "jpeg: bad DHT segment");
        const int cls = s[q] >> 4, tid = s[q] & 15;
        int ns = 0;
        for (int i = 0; i < 16; ++i) ns += s[q + 1 + i];
        SAG_REQUIRE(cls < 2 && tid < 4 && ns <= 256 && q + 17 + ns <= sl, SAG_EINVAL, "jpeg: bad DHT segment");
        SAG_REQUIRE(build_table(cls ? h->ac[tid] : h->dc[tid], s + q + 1, s + q + 17, ns) == SAG_OK, SAG_EINVAL,
                    "jpeg: inconsistent Huffman table");
        q += 17 + ns;
      }
    } else if (m == 0xE0) {
      if (sl >= 5 && memcmp(s, "JFIF", 5) == 0) saw_jfif = true;
    } else if (m == 0xEE) {
      if (sl >= 12 && memcmp(s, "Adobe", 5) == 0) adobe_transform = s[11];
    } else if (m == 0xDD) {
      SAG_REQUIRE(sl >= 2, SAG_EINVAL, "jpeg: bad DRI segment");
      h->restart_interval = (s[0] << 8) | s[1];
    } else if (m == 0xDA) {
      SAG_REQUIRE(have_sof, SAG_EINVAL, "jpeg: scan before the frame header");
      SAG_REQUIRE(sl >= 1 && s[0] == h->ncomp && sl >= (size_t)1 + 2 * h->ncomp, SAG_EUNSUPPORTED,
                  "jpeg: only one interleaved scan holding every component is supported");
      for (int i = 0; i < h->ncomp; ++i) {
        SAG_REQUIRE(s[1 + 2 * i] == h->comp[i].id, SAG_EUNSUPPORTED, "jpeg: scan components out of frame order");
        h->comp[i].td = s[2 + 2 * i] >> 4;
        h->comp[i].ta = s[2 + 2 * i] & 15;
        SAG_REQUIRE(h->comp[i].td < 4 && h->comp[i].ta < 4 && h->dc[h->comp[i].td].present && h->ac[h->comp[i].ta].present &&
                        h->qt_present[h->comp[i].tq],
                    SAG_EINVAL, "jpeg: scan refers to a table the file does not define");
      }
      h->ecs = d + p;
      h->ecs_size = n - p;
      SAG_REQUIRE(h->ecs_size < ((size_t)1 << 27), SAG_EUNSUPPORTED, "jpeg: scans of 128 MB or more are not supported (32-bit bit positions)");
      break;
    }
  }
  SAG_REQUIRE(h->width > 0 && h->height > 0, SAG_EINVAL, "jpeg: empty image");
  // colour space as libjpeg guesses it (jdapimin.c default_decompress_parms): three components are YCbCr unless an Adobe marker
  // says "no transform" or, without JFIF / Adobe markers, the component ids spell R, G, B -- those files hold RGB samples
  if (h->ncomp == 3 && !saw_jfif) {
    const bool rgb_ids = h->comp[0].id == 'R' && h->comp[1].id == 'G' && h->comp[2].id == 'B';
    SAG_REQUIRE(!(adobe_transform == 0 || (adobe_transform < 0 && rgb_ids)), SAG_EUNSUPPORTED,
                "jpeg: the file stores RGB samples (Adobe transform 0 / component ids R, G, B): only YCbCr and grey files are supported");
  }
  if (h->ncomp == 1) h->comp[0].h = h->comp[0].v = 1;          // a one-component scan is never interleaved: MCU = one block
  h->hmax = h->vmax = 1;
  for (int i = 0; i < h->ncomp; ++i) { h->hmax = std::max(h->hmax, h->comp[i].h); h->vmax = std::max(h->vmax, h->comp[i].v); }
  for (int i = 0; i < h->ncomp; ++i) {
    const int rh = h->hmax / h->comp[i].h, rv = h->vmax / h->comp[i].v;
    const bool ok = rh * h->comp[i].h == h->hmax && rv * h->comp[i].v == h->vmax &&
                    ((rh == 1 && rv == 1) || (rh == 2 && rv == 1) || (rh == 2 && rv == 2));
    SAG_REQUIRE(ok && h->comp[i].h <= 2 && h->comp[i].v <= 2, SAG_EUNSUPPORTED,
                "jpeg: component %d sampled %dx%d of %dx%d (supported: full, 2:1 horizontal, 2:1 both)", i, h->comp[i].h, h->comp[i].v,
                h->hmax, h->vmax);
  }
  int blocks_per_mcu = 0;
  for (int i = 0; i < h->ncomp; ++i) blocks_per_mcu += h->comp[i].h * h->comp[i].v;
  SAG_REQUIRE(blocks_per_mcu <= 10, SAG_EINVAL, "jpeg: %d blocks per MCU (the standard allows 10)", blocks_per_mcu);
  h->mcux = (h->width + 8 * h->hmax - 1) / (8 * h->hmax);
  h->mcuy = (h->height + 8 * h->vmax - 1) / (8 * h->vmax);
  for (int i = 0; i < h->ncomp; ++i) { h->bw[i] = h->mcux * h->comp[i].h; h->bh[i] = h->mcuy * h->comp[i].v; }
  return SAG_OK;
}

// ---- entropy decoder (jdhuff.c decode_mcu) ---------------------------------------------------------------------------
struct BitReader {
  const uint8_t* d;
  size_t n, p = 0;
  uint64_t acc = 0;
  int bits = 0;
  BitReader(const uint8_t* d_, size_t n_) : d(d_), n(n_) {}
  inline void fill() {                       // top up to at least 32 valid bits; past a marker the stream reads as zeros
    while (bits <= 56) {
      uint64_t b = 0;
      if (p < n) {
        b = d[p];
        if (b == 0xFF) {
          const int nxt = p + 1 < n ? d[p + 1] : 0xD9;
          if (nxt == 0) p += 2; else b = 0;  // stuffed byte / marker: stay on it
        } else {
          ++p;
        }
      }
      acc = (acc << 8) | b;
      bits += 8;
    }
  }
  inline uint32_t peek(int k) { return (uint32_t)((acc >> (bits - k)) & ((1u << k) - 1)); }
  inline void skip(int k) { bits -= k; }
  inline uint32_t get(int k) { const uint32_t v = peek(k); bits -= k; return v; }
  void restart() {                           // discard the partial byte, skip the RSTn marker
    acc = 0;
    bits = 0;
    while (p + 1 < n && !(d[p] == 0xFF && d[p + 1] >= 0xD0 && d[p + 1] <= 0xD7)) ++p;
    p += 2;
  }
};

inline int huff_decode(BitReader& br, const HuffTable& t) {
  if (br.bits < 32) br.fill();
  const uint16_t e = t.look[br.peek(9)];
  if (e) { br.skip(e >> 8); return e & 255; }
  int len = 10;
  int32_t code = (int32_t)br.peek(10);
  while (len <= 16 && code > t.maxcode[len]) { ++len; code = (int32_t)br.peek(len); }
  if (len > 16) return -1;
  br.skip(len);
  return t.symbols[(code + t.valoffset[len]) & 255];
}

inline int extend(int r, int s) { return r < (1 << (s - 1)) ? r - (1 << s) + 1 : r; }   // jdhuff.h HUFF_EXTEND

// coef[c]: bh[c] x bw[c] blocks of 64 int16 in natural order, zero-initialised by the caller
int decode_scan(const Header& h, int16_t* const coef[3]) {
  BitReader br(h.ecs, h.ecs_size);
  int pred[3] = {0, 0, 0};
  int left = h.restart_interval;
  for (int my = 0; my < h.mcuy; ++my) {
    for (int mx = 0; mx < h.mcux; ++mx) {
      if (h.restart_interval && left == 0) {
        br.restart();
        pred[0] = pred[1] = pred[2] = 0;
        left = h.restart_interval;
      }
      for (int c = 0; c < h.ncomp; ++c) {
        const Component& cp = h.comp[c];
        const HuffTable& dc = h.dc[cp.td];
        const HuffTable& ac = h.ac[cp.ta];
        for (int by = 0; by < cp.v; ++by) {
          for (int bx = 0; bx < cp.h; ++bx) {
            int16_t* blk = coef[c] + ((size_t)(my * cp.v + by) * h.bw[c] + (mx * cp.h + bx)) * 64;
            int s = huff_decode(br, dc);
            if (s < 0 || s > 15) { set_error("jpeg: corrupt entropy-coded data (DC)"); return SAG_EINVAL; }
            if (s) { br.fill(); pred[c] += extend((int)br.get(s), s); }
            blk[0] = (int16_t)pred[c];
            for (int k = 1; k < 64; ++k) {
              const int rs = huff_decode(br, ac);
              if (rs < 0) { set_error("jpeg: corrupt entropy-coded data (AC)"); return SAG_EINVAL; }
              const int r = rs >> 4;
              s = rs & 15;
              if (s) {
                k += r;
                if (br.bits < 32) br.fill();
                blk[kZigzag[k & 63]] = (int16_t)extend((int)br.get(s), s);
              } else if (r == 15) {
                k += 15;
              } else {
                break;
              }
            }
          }
        }
      }
      --left;
    }
  }
  return SAG_OK;
}

// ---- device side --------------------------------------------------------------------------------------------------------
struct JpegImage {            // one frame of a batch (device copy)
  int ncomp, hmax, vmax;
  int h[3], v[3], bw[3], bh[3];
  long long coef_off[3];      // int16 elements into the coefficient buffer
  long long plane_off[3];     // bytes into the plane buffer; row stride bw * 8
  long long block_base[4];    // running count of the blocks of components 0..2 (block_base[3] = all)
};

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// one pass of jpeg_idct_islow over eight values (jidctint.c; CONST_BITS 13)
__device__ __forceinline__ void islow_1d(const int* in, int* out, int shift) {
  int z2 = in[2], z3 = in[6];
  int z1 = (z2 + z3) * 4433;
  int tmp2 = z1 + z3 * (-15137);
  int tmp3 = z1 + z2 * 6270;
  z2 = in[0];
  z3 = in[4];
  int tmp0 = (z2 + z3) << 13, tmp1 = (z2 - z3) << 13;
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  tmp0 = in[7];
  tmp1 = in[5];
  tmp2 = in[3];
  tmp3 = in[1];
  z1 = tmp0 + tmp3;
  z2 = tmp1 + tmp2;
  z3 = tmp0 + tmp2;
  int z4 = tmp1 + tmp3;
  const int z5 = (z3 + z4) * 9633;
  tmp0 *= 2446;
  tmp1 *= 16819;
  tmp2 *= 25172;
  tmp3 *= 12299;
  z1 *= -7373;
  z2 *= -20995;
  z3 = z3 * (-16069) + z5;
  z4 = z4 * (-3196) + z5;
  tmp0 += z1 + z3;
  tmp1 += z2 + z4;
  tmp2 += z2 + z3;
  tmp3 += z1 + z4;
  out[0] = descale(tmp10 + tmp3, shift);
  out[7] = descale(tmp10 - tmp3, shift);
  out[1] = descale(tmp11 + tmp2, shift);
  out[6] = descale(tmp11 - tmp2, shift);
  out[2] = descale(tmp12 + tmp1, shift);
  out[5] = descale(tmp12 - tmp1, shift);
  out[3] = descale(tmp13 + tmp0, shift);
  out[4] = descale(tmp13 - tmp0, shift);
}

// grid (ceil(blocks of the largest frame / 32), n): 8 threads per 8x8 block -- thread j transforms column j, then row j
constexpr int kIdctBlocksPerCta = 32;
__global__ void __launch_bounds__(256) jpeg_idct_kernel(const JpegImage* __restrict__ images, const int16_t* __restrict__ coef,
                                                        const uint16_t* __restrict__ qt, uint8_t* __restrict__ planes) {
  __shared__ int ws[kIdctBlocksPerCta][64 + 8];        // (+8: rows of 9 words, conflict-free column writes)
  const JpegImage& im = images[blockIdx.y];
  const int lb = threadIdx.x >> 3, j = threadIdx.x & 7;
  const long long b = (long long)blockIdx.x * kIdctBlocksPerCta + lb;
  const bool live = b < im.block_base[3];
  int c = 0;
  if (live) { while (b >= im.block_base[c + 1]) ++c; }
  const long long bi = live ? b - im.block_base[c] : 0;
  int v[8], o[8];
  if (live) {
    const int16_t* src = coef + im.coef_off[c] + bi * 64;
    const uint16_t* q = qt + ((long long)blockIdx.y * 3 + c) * 64;
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = (int)src[r * 8 + j] * (int)q[r * 8 + j];       // column j, dequantised
    islow_1d(v, o, 13 - 2);
#pragma unroll
    for (int r = 0; r < 8; ++r) ws[lb][r * 9 + j] = o[r];
  }
  __syncthreads();
  if (live) {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = ws[lb][j * 9 + k];                              // row j of the workspace
    islow_1d(v, o, 13 + 2 + 3);
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int x = (((o[k] & 1023) ^ 512) - 512) + 128;                                     // range_limit[x & RANGE_MASK]
      x = min(max(x, 0), 255);
      if (k < 4) lo |= (uint32_t)x << (8 * k); else hi |= (uint32_t)x << (8 * (k - 4));
    }
    const int bx = (int)(bi % im.bw[c]), by = (int)(bi / im.bw[c]);
    uint8_t* dst = planes + im.plane_off[c] + ((long long)(by * 8 + j) * im.bw[c] + bx) * 8;
    *reinterpret_cast<uint2*>(dst) = make_uint2(lo, hi);
  }
}

// sample of component c at full resolution (jdsample.c fancy upsamplers)
__device__ __forceinline__ int jpeg_sample(const JpegImage& im, const uint8_t* __restrict__ planes, int c, int x, int y, int width,
                                           int height) {
  const uint8_t* p = planes + im.plane_off[c];
  const int stride = im.bw[c] * 8;
  const int rh = im.hmax / im.h[c], rv = im.vmax / im.v[c];
  if (rh == 1) return p[(long long)y * stride + x];
  const int dw = (width * im.h[c] + im.hmax - 1) / im.hmax;     // downsampled_width: the real samples of a row
  const int cx = x >> 1;
  // components at most two samples wide are replicated, not filtered (jdsample.c jinit_upsampler: fancy only if downsampled_width > 2)
  if (dw <= 2) return p[(long long)(rv == 1 ? y : (y >> 1)) * stride + cx];
  if (rv == 1) {                                                // h2v1_fancy_upsample
    const uint8_t* r = p + (long long)y * stride;
    const int s = r[cx];
    if (x & 1) return cx == dw - 1 ? s : (3 * s + r[cx + 1] + 2) >> 2;
    return cx == 0 ? s : (3 * s + r[cx - 1] + 1) >> 2;
  }
  const int dh = (height * im.v[c] + im.vmax - 1) / im.vmax;    // h2v2_fancy_upsample
  const int cy = y >> 1;
  const int ny = (y & 1) ? min(cy + 1, dh - 1) : max(cy - 1, 0);
  const uint8_t* r0 = p + (long long)cy * stride;
  const uint8_t* r1 = p + (long long)ny * stride;
  const int s = 3 * r0[cx] + r1[cx];
  if (x & 1) return cx == dw - 1 ? (4 * s + 7) >> 4 : (3 * s + 3 * r0[cx + 1] + r1[cx + 1] + 7) >> 4;
  return cx == 0 ? (4 * s + 8) >> 4 : (3 * s + 3 * r0[cx - 1] + r1[cx - 1] + 8) >> 4;
}

// grid (ceil(W / 4 / 64), H, n): four pixels (12 bytes) per thread
__global__ void __launch_bounds__(64) jpeg_rgb_kernel(const JpegImage* __restrict__ images, const uint8_t* __restrict__ planes, int width,
                                                      int height, uint8_t* __restrict__ frames) {
  const JpegImage& im = images[blockIdx.z];
  const int y = blockIdx.y;
  const int x0 = (blockIdx.x * 64 + threadIdx.x) * 4;
  if (x0 >= width) return;
  uint8_t px[12];
  const int cnt = min(4, width - x0);
  for (int i = 0; i < cnt; ++i) {
    const int x = x0 + i;
    const int yy = jpeg_sample(im, planes, 0, x, y, width, height);
    int r = yy, g = yy, b = yy;
    if (im.ncomp == 3) {                                        // jdcolor.c ycc_rgb_convert (SCALEBITS 16)
      const int cb = jpeg_sample(im, planes, 1, x, y, width, height) - 128;
      const int cr = jpeg_sample(im, planes, 2, x, y, width, height) - 128;
      r = yy + ((91881 * cr + 32768) >> 16);
      g = yy + ((-22554 * cb + 32768 - 46802 * cr) >> 16);
      b = yy + ((116130 * cb + 32768) >> 16);
    }
    px[3 * i + 0] = (uint8_t)min(max(r, 0), 255);
    px[3 * i + 1] = (uint8_t)min(max(g, 0), 255);
    px[3 * i + 2] = (uint8_t)min(max(b, 0), 255);
  }
  uint8_t* dst = frames + (((long long)blockIdx.z * height + y) * width + x0) * 3;
  if (cnt == 4 && ((reinterpret_cast<uintptr_t>(dst) & 3) == 0)) {
    uint32_t w[3];
    memcpy(w, px, 12);
    reinterpret_cast<uint32_t*>(dst)[0] = w[0];
    reinterpret_cast<uint32_t*>(dst)[1] = w[1];
    reinterpret_cast<uint32_t*>(dst)[2] = w[2];
  } else {
    for (int i = 0; i < 3 * cnt; ++i) dst[i] = px[i];
  }
}


// ---- parallel entropy decoding on the device --------------------------------------------------------------------------------
// A Huffman-coded scan is one serial bit stream, but its decoder resynchronises by itself: started at an arbitrary bit with a
// guessed state it falls into step with the true decoder after a few symbols.  The unstuffed stream of each frame is cut
// into subsequences of a few hundred bits, one per thread (Weissenberger & Schmidt's scheme):
//   round 0   every subsequence is decoded from its first bit as if a block started there; the state at its end -- (bit
//             position of the next symbol, block of the MCU, zig-zag index) -- is recorded
//   round r   every subsequence is decoded again from the recorded end state of its predecessor, if that changed; the first
//             subsequence of a restart segment always starts from the true state, so after r rounds the first r subsequences
//             are exact, and in practice two or three rounds make every state consistent -- which is the stopping rule
//   scan      blocks completed per subsequence -> index of the block each subsequence starts in
//   write     a last decode stores the coefficients (DC as differences) at their places in the frame's block grids
//   DC        prefix sums of the DC differences per component in scan order, restarted at every restart segment
// All phases are __host__ __device__ functions of (thread, threads): the kernel runs them with one CTA per frame and
// __syncthreads between phases; sag_jpeg_coefficients_parallel runs the same code thread by thread on the host (tests).
struct SubState { int p; short b, z; };
__host__ __device__ inline bool same_state(const SubState& a, const SubState& b) { return a.p == b.p && a.b == b.b && a.z == b.z; }

struct JpegSub {              // one subsequence
  int begin_bit, end_bit;     // in the frame's unstuffed stream
  int first;                  // starts a restart segment: the true initial state is known
  int seg_block0, seg_block1; // scan-order indices of the segment's first block and of the one past its last
};

struct JpegStream {           // the scan of one frame
  long long data_off;         // byte offset of the unstuffed stream in the batch's buffer (padded with 16 zero bytes)
  int sub0, n_subs;           // its rows of the batch's JpegSub table
  int bpm;                    // blocks per MCU
  unsigned comp_nib;          // component of block i of an MCU in bits 2i, 2i+1 (what the decode loop reads)
  int blk_comp[10], blk_x[10], blk_y[10];   // (at most 10 blocks per MCU: the standard's limit, checked by parse_header)
  int mcux, n_mcus;
  int restart_mcus;           // MCUs per restart segment (0: one segment)
  int dc_tab[3], ac_tab[3];   // rows of the batch's table array
};

// the frame's tables, contiguous: DC of components 0..ncomp-1, then AC (addresses by arithmetic: no pointer array in local memory)
struct HuffView { const HuffTable* base; int ncomp; };

// The stream as big-endian 32-bit words with the two words around the read position cached in registers: a symbol costs a
// global load only when the position crosses a word (streams start 16-byte aligned and are padded with 16 zero bytes).
struct BitWindow {
  const uint32_t* words;
  int idx;
  uint32_t w0, w1;
  __host__ __device__ static inline uint32_t be(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __byte_perm(x, 0, 0x0123);
#else
    return __builtin_bswap32(x);
#endif
  }
  __host__ __device__ inline void init(const uint8_t* s, int p) {
    words = reinterpret_cast<const uint32_t*>(s);
    idx = p >> 5;
    w0 = be(words[idx]);
    w1 = be(words[idx + 1]);
  }
  __host__ __device__ inline uint32_t peek32(int p) {     // the 32 bits that start at bit p (p never moves backwards)
    const int i = p >> 5;
    if (i != idx) {
      if (i == idx + 1) { w0 = w1; w1 = be(words[i + 1]); }
      else { w0 = be(words[i]); w1 = be(words[i + 1]); }
      idx = i;
    }
#ifdef __CUDA_ARCH__
    return __funnelshift_l(w1, w0, p & 31);             // high word of (w0:w1) << (p % 32): one instruction
#else
    return (uint32_t)(((((uint64_t)w0) << 32) | (uint64_t)w1) >> (32 - (p & 31)));
#endif
  }
};

// one Huffman symbol at the top of `w`: its code length (>= 1, so the decoder always advances) and value
__host__ __device__ inline int huff_symbol(const HuffTable& t, uint32_t w, int* len) {
  const uint16_t e = t.look[w >> 23];
  if (e) { *len = e >> 8; return e & 255; }
  for (int l = 10; l <= 16; ++l) {
    const int32_t code = (int32_t)(w >> (32 - l));
    if (code <= t.maxcode[l]) { *len = l; return t.symbols[(code + t.valoffset[l]) & 255]; }
  }
  *len = 16;                  // no such code (only off the true path, or in a corrupt file)
  return 0;
}

__host__ __device__ inline int16_t* block_ptr(const JpegStream& js, const JpegImage& im, int16_t* coef, long long q) {
  const int m = (int)(q / js.bpm), bi = (int)(q % js.bpm);
  const int c = js.blk_comp[bi];
  const int bx = (m % js.mcux) * im.h[c] + js.blk_x[bi], by = (m / js.mcux) * im.v[c] + js.blk_y[bi];
  return coef + im.coef_off[c] + ((long long)by * im.bw[c] + bx) * 64;
}

// Decodes from state `st` up to bit `end_bit`; returns the number of blocks completed.  WRITE: q = scan-order index of the block
// the state is in, q_end = one past the segment's last block; coefficients go to their places (DC as the difference).
template <bool WRITE>
__host__ __device__ inline int decode_span(const uint8_t* s, const HuffView& hv, const JpegStream& js, SubState& st, int end_bit,
                                           long long q, long long q_end, const JpegImage* im, int16_t* coef, const uint8_t* zigzag) {
  int p = st.p, b = st.b, z = st.z, done = 0;
  const int bpm = js.bpm, ncomp = hv.ncomp;             // (loop invariants in registers: js lives in shared memory)
  const unsigned nib = js.comp_nib;
  const HuffTable* const tabs = hv.base;
  BitWindow bw;
  bw.init(s, p);
  int16_t* blk = nullptr;
  if (WRITE && q < q_end) blk = block_ptr(js, *im, coef, q);
  while (p < end_bit) {
    if (WRITE && q >= q_end) break;                     // only padding bits are left in this restart segment
    const int c = (int)((nib >> (2 * b)) & 3u);
    const uint32_t w = bw.peek32(p);
    // one path for DC and AC symbols (threads of a warp sit at different places of their blocks): a DC symbol is its size
    // (<= 11, so its "run" nibble is 0), an AC symbol (run << 4) | size
    int len;
    const int rs = huff_symbol(tabs[z == 0 ? c : ncomp + c], w, &len);
    const int r = rs >> 4, sz = rs & 15;
    p += len + sz;
    if (sz) {
      z += r;
      if (WRITE) {
        const int v = (int)((w << len) >> (32 - sz));
        blk[zigzag[z & 63]] = (int16_t)(v < (1 << (sz - 1)) ? v - (1 << sz) + 1 : v);
      }
      z += 1;
    } else if (z == 0) {
      z = 1;                      // DC difference 0 (the block is zero-initialised)
    } else if (r == 15) {
      z += 16;
    } else {
      z = 64;
    }
    if (z >= 64) {
      z = 0;
      b = b + 1 == bpm ? 0 : b + 1;
      ++done;
      if (WRITE) { ++q; if (q < q_end) blk = block_ptr(js, *im, coef, q); }
    }
  }
  st.p = p;
  st.b = (short)b;
  st.z = (short)z;
  return done;
}

struct HuffScratch {          // per subsequence of the batch (device memory)
  SubState* exit[2];
  SubState* entry;
  int* nblk;
  int* base;
};

// one synchronisation round over the frame's subsequences; returns whether this thread re-decoded anything
__host__ __device__ inline int phase_sync(int tid, int nthreads, int round, int cur, const uint8_t* s, const HuffView& hv, const JpegStream& js,
                                          const JpegSub* subs, const HuffScratch& sc) {
  int changed = 0;
  for (int k = tid; k < js.n_subs; k += nthreads) {
    const int j = js.sub0 + k;
    const JpegSub& sb = subs[j];
    SubState e;
    if (sb.first || round == 0) { e.p = sb.begin_bit; e.b = 0; e.z = 0; }
    else e = sc.exit[cur][j - 1];
    if (round > 0 && same_state(e, sc.entry[j])) { sc.exit[cur ^ 1][j] = sc.exit[cur][j]; continue; }
    sc.entry[j] = e;
    sc.nblk[j] = decode_span<false>(s, hv, js, e, sb.end_bit, 0, 0, nullptr, nullptr, nullptr);
    sc.exit[cur ^ 1][j] = e;
    changed = 1;
  }
  return changed;
}

// segmented exclusive scan of nblk -> base, in three steps (local, carries by thread 0, apply); total / reset: nthreads ints each
__host__ __device__ inline void phase_scan_local(int tid, int nthreads, const JpegStream& js, const JpegSub* subs, const HuffScratch& sc, int* total,
                                                 int* reset) {
  const int per = (js.n_subs + nthreads - 1) / nthreads;
  int run = 0, has_reset = 0;
  for (int k = tid * per; k < min((tid + 1) * per, js.n_subs); ++k) {
    const int j = js.sub0 + k;
    if (subs[j].first) { run = subs[j].seg_block0; has_reset = 1; }
    sc.base[j] = run;
    run += sc.nblk[j];
  }
  total[tid] = run;
  reset[tid] = has_reset;
}
__host__ __device__ inline void phase_scan_carry(int nthreads, int* total, const int* reset) {     // total[t] <- carry into thread t
  int carry = 0;
  for (int t = 0; t < nthreads; ++t) {
    const int mine = total[t];
    total[t] = carry;
    carry = reset[t] ? mine : carry + mine;
  }
}
__host__ __device__ inline void phase_scan_apply(int tid, int nthreads, const JpegStream& js, const JpegSub* subs, const HuffScratch& sc, const int* total) {
  const int per = (js.n_subs + nthreads - 1) / nthreads;
  for (int k = tid * per; k < min((tid + 1) * per, js.n_subs); ++k) {
    const int j = js.sub0 + k;
    if (subs[j].first) break;
    sc.base[j] += total[tid];
  }
}

__host__ __device__ inline void phase_write(int tid, int nthreads, const uint8_t* s, const HuffView& hv, const JpegStream& js, const JpegSub* subs,
                                            const HuffScratch& sc, const JpegImage& im, int16_t* coef, const uint8_t* zigzag) {
  for (int k = tid; k < js.n_subs; k += nthreads) {
    const int j = js.sub0 + k;
    SubState e = sc.entry[j];
    decode_span<true>(s, hv, js, e, subs[j].end_bit, sc.base[j], subs[j].seg_block1, &im, coef, zigzag);
  }
}

// DC prediction of component c: inclusive sums of the differences in scan order, restarted every `seg` blocks (0: never)
__host__ __device__ inline int16_t* dc_block(const JpegStream& js, const JpegImage& im, int16_t* coef, int c, int k) {
  const int nb = im.h[c] * im.v[c];
  const int m = k / nb, i = k % nb;
  const int bx = (m % js.mcux) * im.h[c] + i % im.h[c], by = (m / js.mcux) * im.v[c] + i / im.h[c];
  return coef + im.coef_off[c] + ((long long)by * im.bw[c] + bx) * 64;
}
__host__ __device__ inline void phase_dc_local(int tid, int nthreads, const JpegStream& js, const JpegImage& im, int16_t* coef, int c, int* total,
                                               int* reset) {
  const int nb = im.h[c] * im.v[c], n = js.n_mcus * nb, seg = js.restart_mcus * nb;
  const int per = (n + nthreads - 1) / nthreads;
  int run = 0, has_reset = 0;
  for (int k = tid * per; k < min((tid + 1) * per, n); ++k) {
    if (seg > 0 && k % seg == 0) { run = 0; has_reset = 1; }
    int16_t* d = dc_block(js, im, coef, c, k);
    run += *d;
    *d = (int16_t)run;
  }
  total[tid] = run;
  reset[tid] = has_reset;
}
__host__ __device__ inline void phase_dc_apply(int tid, int nthreads, const JpegStream& js, const JpegImage& im, int16_t* coef, int c, const int* total) {
  const int nb = im.h[c] * im.v[c], n = js.n_mcus * nb, seg = js.restart_mcus * nb;
  const int per = (n + nthreads - 1) / nthreads;
  if (total[tid] == 0) return;
  for (int k = tid * per; k < min((tid + 1) * per, n); ++k) {
    if (seg > 0 && k % seg == 0) break;
    int16_t* d = dc_block(js, im, coef, c, k);
    *d = (int16_t)(*d + total[tid]);
  }
}

__constant__ uint8_t c_zigzag[64];

constexpr int kHuffThreads = 256;
// grid (n frames): one CTA per frame
__global__ void __launch_bounds__(kHuffThreads) jpeg_huffman_kernel(const JpegStream* __restrict__ streams, const JpegSub* __restrict__ subs,
                                                                    const HuffTable* __restrict__ tables, const JpegImage* __restrict__ images,
                                                                    const uint8_t* __restrict__ data, HuffScratch sc, int16_t* __restrict__ coef,
                                                                    int* __restrict__ rounds_out) {
  __shared__ __align__(16) unsigned char tab_raw[6 * sizeof(HuffTable)];
  __shared__ int total[kHuffThreads], reset[kHuffThreads];
  __shared__ JpegStream js;
  __shared__ JpegImage im;
  const int tid = threadIdx.x;
  if (tid == 0) { js = streams[blockIdx.x]; im = images[blockIdx.x]; }
  __syncthreads();
  HuffTable* tab = reinterpret_cast<HuffTable*>(tab_raw);
  for (int t = 0; t < 2 * im.ncomp; ++t) {               // the frame's tables into shared memory (word copies)
    const int row = t < im.ncomp ? js.dc_tab[t] : js.ac_tab[t - im.ncomp];
    const uint32_t* src = reinterpret_cast<const uint32_t*>(tables + row);
    uint32_t* dst = reinterpret_cast<uint32_t*>(tab + t);
    for (int i = tid; i < (int)(sizeof(HuffTable) / 4); i += kHuffThreads) dst[i] = src[i];
  }
  __syncthreads();
  HuffView hv;
  hv.base = tab;
  hv.ncomp = im.ncomp;
  const uint8_t* s = data + js.data_off;
  int cur = 0, round = 0;
  for (;; ++round) {
    const int changed = phase_sync(tid, kHuffThreads, round, cur, s, hv, js, subs, sc);
    cur ^= 1;
    if (!__syncthreads_or(changed)) break;
  }
  if (tid == 0 && rounds_out != nullptr) rounds_out[blockIdx.x] = round;
  phase_scan_local(tid, kHuffThreads, js, subs, sc, total, reset);
  __syncthreads();
  if (tid == 0) phase_scan_carry(kHuffThreads, total, reset);
  __syncthreads();
  phase_scan_apply(tid, kHuffThreads, js, subs, sc, total);
  __syncthreads();
  phase_write(tid, kHuffThreads, s, hv, js, subs, sc, im, coef, c_zigzag);
  __syncthreads();
  for (int c = 0; c < im.ncomp; ++c) {
    phase_dc_local(tid, kHuffThreads, js, im, coef, c, total, reset);
    __syncthreads();
    if (tid == 0) phase_scan_carry(kHuffThreads, total, reset);
    __syncthreads();
    phase_dc_apply(tid, kHuffThreads, js, im, coef, c, total);
    __syncthreads();
  }
}

// ---- host preparation of a scan for the parallel decoder -------------------------------------------------------------------
// Unstuffs the entropy-coded segment (FF 00 -> FF), drops the RSTn markers and records where each restart segment starts.
void unstuff(const Header& h, std::vector<uint8_t>& out, std::vector<int>& seg_start) {
  out.clear();
  seg_start.assign(1, 0);
  const uint8_t* d = h.ecs;
  const size_t n = h.ecs_size;
  out.reserve(n + 16);
  size_t p = 0;
  while (p < n) {
    const uint8_t* f = static_cast<const uint8_t*>(memchr(d + p, 0xFF, n - p));
    const size_t q = f ? (size_t)(f - d) : n;
    out.insert(out.end(), d + p, d + q);
    if (!f) break;
    const int nxt = q + 1 < n ? d[q + 1] : 0xD9;
    if (nxt == 0) { out.push_back(0xFF); p = q + 2; }
    else if (nxt >= 0xD0 && nxt <= 0xD7) { seg_start.push_back((int)out.size()); p = q + 2; }
    else if (nxt == 0xFF) { p = q + 1; }
    else break;                                           // EOI or another marker: the scan ends here
  }
}

int g_min_sub_bytes = 256;     // measured on B200, 32 frames of 224x448 at quality 90: 64 B -> 11-22 rounds, 256 B -> 3-6, fastest
int sub_bytes_for(size_t stream_bytes) {                  // at most 2048 subsequences per frame, at least g_min_sub_bytes each
  size_t sb = (stream_bytes + 2047) / 2048;
  sb = std::max<size_t>((size_t)g_min_sub_bytes, (sb + 3) / 4 * 4);
  return (int)sb;
}

// fills js (except data_off / sub0 / table rows) and appends the frame's subsequences
void plan_stream(const Header& h, const std::vector<int>& seg_start, int stream_bytes, JpegStream* js, std::vector<JpegSub>* subs) {
  memset(js, 0, sizeof(*js));
  js->bpm = 0;
  for (int c = 0; c < h.ncomp; ++c)
    for (int y = 0; y < h.comp[c].v; ++y)
      for (int x = 0; x < h.comp[c].h; ++x) {
        js->blk_comp[js->bpm] = c; js->blk_x[js->bpm] = x; js->blk_y[js->bpm] = y;
        js->comp_nib |= (unsigned)c << (2 * js->bpm);
        ++js->bpm;
      }
  js->mcux = h.mcux;
  js->n_mcus = h.mcux * h.mcuy;
  js->restart_mcus = h.restart_interval;
  const int total_blocks = js->n_mcus * js->bpm;
  const int seg_blocks = h.restart_interval ? h.restart_interval * js->bpm : total_blocks;
  const int sb = sub_bytes_for((size_t)stream_bytes);
  const int first_row = (int)subs->size();
  for (size_t g = 0; g < seg_start.size(); ++g) {
    const int b0 = seg_start[g], b1 = g + 1 < seg_start.size() ? seg_start[g + 1] : stream_bytes;
    const long long q0 = (long long)g * seg_blocks;
    if (q0 >= total_blocks) break;                        // (markers past the last MCU)
    for (int o = b0; o < b1 || o == b0; o += sb) {
      JpegSub s;
      s.begin_bit = o * 8;
      s.end_bit = std::min(o + sb, std::max(b1, b0)) * 8;
      s.first = o == b0;
      s.seg_block0 = (int)q0;
      s.seg_block1 = (int)std::min<long long>(q0 + seg_blocks, total_blocks);
      subs->push_back(s);
      if (b1 <= b0) break;
    }
  }
  js->n_subs = (int)subs->size() - first_row;
}

}  // namespace
}  // namespace sag

struct sag_jpeg {
  int max_frames = 0, height = 0, width = 0, device = 0;
  size_t coef_cap = 0, plane_cap = 0;      // per frame: int16 elements / bytes (4:4:4 worst case, whole 16x16 MCUs)
  int16_t* h_coef = nullptr;               // pinned staging
  uint16_t* h_qt = nullptr;
  sag::JpegImage* h_img = nullptr;
  int16_t* d_coef = nullptr;
  uint16_t* d_qt = nullptr;
  sag::JpegImage* d_img = nullptr;
  uint8_t* d_planes = nullptr;
  cudaEvent_t staged = nullptr;            // the previous call's host -> device copies have left the staging buffers
  bool staged_pending = false;
  // device entropy decoding (option "device_huffman", default on): the unstuffed scans travel instead of the coefficients
  int device_huffman = 1;
  size_t stream_cap = 0, sub_cap = 0, table_cap = 0;
  uint8_t* h_stream = nullptr;             // pinned
  uint8_t* d_stream = nullptr;
  sag::JpegSub* h_subs = nullptr;          // pinned
  sag::JpegSub* d_subs = nullptr;
  sag::JpegStream* h_js = nullptr;         // pinned, max_frames
  sag::JpegStream* d_js = nullptr;
  sag::HuffTable* h_tables = nullptr;      // pinned
  sag::HuffTable* d_tables = nullptr;
  char* d_scratch = nullptr;               // SubState x 3 + int x 2 per subsequence
  int* d_rounds = nullptr;                 // synchronisation rounds of each frame of the last decode
  int last_n = 0;
};

using namespace sag;

extern "C" {

int sag_jpeg_info(const void* host_file, size_t size, int* width, int* height, int* components, int* h_samp, int* v_samp) {
  SAG_REQUIRE(host_file != nullptr, SAG_EINVAL, "jpeg: null file");
  Header* h = new Header();
  const int rc = parse_header(static_cast<const uint8_t*>(host_file), size, h);
  if (rc == SAG_OK) {
    if (width) *width = h->width;
    if (height) *height = h->height;
    if (components) *components = h->ncomp;
    if (h_samp) *h_samp = h->hmax;
    if (v_samp) *v_samp = h->vmax;
  }
  delete h;
  return rc;
}

int sag_jpeg_coefficients(const void* host_file, size_t size, int16_t* host_coef, size_t capacity, int* blocks_wide, int* blocks_high,
                          uint16_t* host_qt) {
  SAG_REQUIRE(host_file != nullptr && host_coef != nullptr, SAG_EINVAL, "jpeg: null argument");
  Header* h = new Header();
  int rc = parse_header(static_cast<const uint8_t*>(host_file), size, h);
  if (rc == SAG_OK) {
    size_t need = 0;
    for (int c = 0; c < h->ncomp; ++c) need += (size_t)h->bw[c] * h->bh[c] * 64;
    if (need > capacity) {
      set_error("jpeg: coefficient buffer holds %zu values, the file needs %zu", capacity, need);
      rc = SAG_ENOMEM;
    } else {
      memset(host_coef, 0, need * sizeof(int16_t));
      int16_t* cp[3] = {nullptr, nullptr, nullptr};
      size_t off = 0;
      for (int c = 0; c < h->ncomp; ++c) { cp[c] = host_coef + off; off += (size_t)h->bw[c] * h->bh[c] * 64; }
      rc = decode_scan(*h, cp);
      for (int c = 0; c < 3; ++c) {
        if (blocks_wide) blocks_wide[c] = c < h->ncomp ? h->bw[c] : 0;
        if (blocks_high) blocks_high[c] = c < h->ncomp ? h->bh[c] : 0;
        if (host_qt && c < h->ncomp) memcpy(host_qt + 64 * c, h->qt[h->comp[c].tq], 128);
      }
    }
  }
  delete h;
  return rc;
}

int sag_jpeg_create(sag_jpeg** out, int max_frames, int height, int width) {
  SAG_REQUIRE(out != nullptr && max_frames > 0 && max_frames < 65536 && height > 0 && width > 0 && height < 65536 && width < 65536,
              SAG_EINVAL, "jpeg: bad decoder geometry (frames, height and width must be in [1, 65535])");
  sag_jpeg* d = new sag_jpeg();
  d->max_frames = max_frames;
  d->height = height;
  d->width = width;
  const size_t hp = ((size_t)height + 15) / 16 * 16, wp = ((size_t)width + 15) / 16 * 16;
  d->coef_cap = 3 * hp * wp;
  d->plane_cap = 3 * hp * wp;
  *out = d;
  cudaError_t e = cudaGetDevice(&d->device);
  if (e == cudaSuccess) e = cudaHostAlloc(&d->h_coef, d->coef_cap * max_frames * sizeof(int16_t), cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaHostAlloc(&d->h_qt, (size_t)max_frames * 3 * 64 * sizeof(uint16_t), cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaHostAlloc(&d->h_img, (size_t)max_frames * sizeof(JpegImage), cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaMalloc(&d->d_coef, d->coef_cap * max_frames * sizeof(int16_t));
  if (e == cudaSuccess) e = cudaMalloc(&d->d_qt, (size_t)max_frames * 3 * 64 * sizeof(uint16_t));
  if (e == cudaSuccess) e = cudaMalloc(&d->d_img, (size_t)max_frames * sizeof(JpegImage));
  if (e == cudaSuccess) e = cudaMalloc(&d->d_planes, d->plane_cap * max_frames);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->staged, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaHostAlloc(&d->h_js, (size_t)max_frames * sizeof(JpegStream), cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaMalloc(&d->d_js, (size_t)max_frames * sizeof(JpegStream));
  if (e == cudaSuccess) e = cudaMalloc(&d->d_rounds, (size_t)max_frames * sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_zigzag, kZigzag, 64);
  if (e != cudaSuccess) {
    set_error("jpeg: %s while creating the decoder: %s", cudaGetErrorName(e), cudaGetErrorString(e));
    sag_jpeg_destroy(d);
    *out = nullptr;
    return SAG_ECUDA;
  }
  return SAG_OK;
}

void sag_jpeg_destroy(sag_jpeg* d) {
  if (d == nullptr) return;
  if (d->staged) { cudaEventSynchronize(d->staged); cudaEventDestroy(d->staged); }
  cudaFreeHost(d->h_coef);
  cudaFreeHost(d->h_qt);
  cudaFreeHost(d->h_img);
  cudaFree(d->d_coef);
  cudaFree(d->d_qt);
  cudaFree(d->d_img);
  cudaFree(d->d_planes);
  cudaFreeHost(d->h_stream);
  cudaFreeHost(d->h_subs);
  cudaFreeHost(d->h_js);
  cudaFreeHost(d->h_tables);
  cudaFree(d->d_stream);
  cudaFree(d->d_subs);
  cudaFree(d->d_js);
  cudaFree(d->d_tables);
  cudaFree(d->d_scratch);
  cudaFree(d->d_rounds);
  delete d;
}

int sag_jpeg_set_option(sag_jpeg* d, const char* key, int value) {
  SAG_REQUIRE(key != nullptr, SAG_EINVAL, "jpeg: null argument");
  if (strcmp(key, "device_huffman") == 0) {
    SAG_REQUIRE(d != nullptr, SAG_EINVAL, "jpeg: device_huffman is an option of a decoder");
    d->device_huffman = value != 0;
    return SAG_OK;
  }
  if (strcmp(key, "sub_bytes") == 0) {                    // (process wide; dec may be NULL) shortest subsequence of the device entropy decoder
    SAG_REQUIRE(value >= 32 && value <= 65536, SAG_EINVAL, "jpeg: sub_bytes must be in [32, 65536]");
    g_min_sub_bytes = value;
    return SAG_OK;
  }
  SAG_REQUIRE(false, SAG_EINVAL, "jpeg: unknown option %s", key);
}

int sag_jpeg_sync_rounds(sag_jpeg* d, int* host_rounds, int n) {
  SAG_REQUIRE(d != nullptr && host_rounds != nullptr && n >= 0 && n <= d->last_n, SAG_EINVAL, "jpeg: the last device decode held %d frames",
              d ? d->last_n : 0);
  SAG_CHECK_CUDA(cudaMemcpy(host_rounds, d->d_rounds, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
  return SAG_OK;
}

}  // extern "C"

namespace {
// pinned + device buffer pair that grows on demand (the caller has synchronised with the copies that read the old one)
template <class T>
int grow(T** host, T** dev, size_t* cap, size_t need) {
  if (need <= *cap) return SAG_OK;
  const size_t want = need + need / 2 + 1024;
  if (*host) SAG_CHECK_CUDA(cudaFreeHost(*host));
  if (*dev) SAG_CHECK_CUDA(cudaFree(*dev));
  *host = nullptr;
  *dev = nullptr;
  *cap = 0;
  SAG_CHECK_CUDA(cudaHostAlloc(reinterpret_cast<void**>(host), want * sizeof(T), cudaHostAllocDefault));
  SAG_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(dev), want * sizeof(T)));
  *cap = want;
  return SAG_OK;
}

HuffScratch carve_scratch(char* base, size_t subs) {
  HuffScratch sc;
  sc.exit[0] = reinterpret_cast<SubState*>(base);
  sc.exit[1] = sc.exit[0] + subs;
  sc.entry = sc.exit[1] + subs;
  sc.nblk = reinterpret_cast<int*>(sc.entry + subs);
  sc.base = sc.nblk + subs;
  return sc;
}
constexpr size_t kScratchPerSub = 3 * sizeof(SubState) + 2 * sizeof(int);
}  // namespace

extern "C" {

// The device algorithm run thread by thread on the host (tests; `nthreads` plays the CTA size): same phase functions, same
// tables, same subsequence plan.  rounds: synchronisation rounds until every state was consistent.
int sag_jpeg_coefficients_parallel(const void* host_file, size_t size, int nthreads, int16_t* host_coef, size_t capacity, int* rounds) {
  SAG_REQUIRE(host_file != nullptr && host_coef != nullptr && nthreads > 0, SAG_EINVAL, "jpeg: bad argument");
  std::unique_ptr<Header> h(new Header());
  SAG_TRY(parse_header(static_cast<const uint8_t*>(host_file), size, h.get()));
  JpegImage im;
  memset(&im, 0, sizeof(im));
  im.ncomp = h->ncomp;
  size_t need = 0;
  for (int c = 0; c < 3; ++c) {
    im.h[c] = im.v[c] = 1;
    if (c < h->ncomp) {
      im.h[c] = h->comp[c].h; im.v[c] = h->comp[c].v; im.bw[c] = h->bw[c]; im.bh[c] = h->bh[c];
      im.coef_off[c] = (long long)need;
      need += (size_t)h->bw[c] * h->bh[c] * 64;
    }
  }
  SAG_REQUIRE(need <= capacity, SAG_ENOMEM, "jpeg: coefficient buffer holds %zu values, the file needs %zu", capacity, need);
  memset(host_coef, 0, need * sizeof(int16_t));
  std::vector<uint8_t> stream;
  std::vector<int> seg;
  unstuff(*h, stream, seg);
  const int nbytes = (int)stream.size();
  stream.resize(stream.size() + 16, 0);
  JpegStream js;
  std::vector<JpegSub> subs;
  plan_stream(*h, seg, nbytes, &js, &subs);
  js.sub0 = 0;
  std::vector<HuffTable> tabs(2 * h->ncomp);
  for (int c = 0; c < h->ncomp; ++c) { tabs[c] = h->dc[h->comp[c].td]; tabs[h->ncomp + c] = h->ac[h->comp[c].ta]; }
  HuffView hv;
  hv.base = tabs.data();
  hv.ncomp = h->ncomp;
  std::vector<char> raw(kScratchPerSub * subs.size() + 64);
  HuffScratch sc = carve_scratch(raw.data(), subs.size());
  std::vector<int> total(nthreads), reset(nthreads);
  int cur = 0, round = 0;
  for (;; ++round) {
    int changed = 0;
    for (int t = 0; t < nthreads; ++t) changed |= phase_sync(t, nthreads, round, cur, stream.data(), hv, js, subs.data(), sc);
    cur ^= 1;
    if (!changed) break;
  }
  if (rounds) *rounds = round;
  for (int t = 0; t < nthreads; ++t) phase_scan_local(t, nthreads, js, subs.data(), sc, total.data(), reset.data());
  phase_scan_carry(nthreads, total.data(), reset.data());
  for (int t = 0; t < nthreads; ++t) phase_scan_apply(t, nthreads, js, subs.data(), sc, total.data());
  for (int t = 0; t < nthreads; ++t) phase_write(t, nthreads, stream.data(), hv, js, subs.data(), sc, im, host_coef, kZigzag);
  for (int c = 0; c < im.ncomp; ++c) {
    for (int t = 0; t < nthreads; ++t) phase_dc_local(t, nthreads, js, im, host_coef, c, total.data(), reset.data());
    phase_scan_carry(nthreads, total.data(), reset.data());
    for (int t = 0; t < nthreads; ++t) phase_dc_apply(t, nthreads, js, im, host_coef, c, total.data());
  }
  return SAG_OK;
}

int sag_jpeg_decode(sag_jpeg* d, const void* const* host_files, const size_t* sizes, int n, uint8_t* frames, int threads, void* stream) {
  SAG_REQUIRE(d != nullptr && host_files != nullptr && sizes != nullptr && frames != nullptr, SAG_EINVAL, "jpeg: null argument");
  SAG_REQUIRE(n > 0 && n <= d->max_frames, SAG_EINVAL, "jpeg: %d frames, the decoder was created for at most %d", n, d->max_frames);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d->staged_pending) {                                     // the staging buffers are about to be overwritten
    SAG_CHECK_CUDA(cudaEventSynchronize(d->staged));
    d->staged_pending = false;
  }
  // headers (serial: microseconds), then the layout of the batch: frames packed back to back
  std::vector<Header> hdr(n);
  size_t coef_total = 0, plane_total = 0;
  long long max_blocks = 0;
  for (int i = 0; i < n; ++i) {
    SAG_REQUIRE(host_files[i] != nullptr, SAG_EINVAL, "jpeg: file %d is null", i);
    SAG_TRY(parse_header(static_cast<const uint8_t*>(host_files[i]), sizes[i], &hdr[i]));
    SAG_REQUIRE(hdr[i].width == d->width && hdr[i].height == d->height, SAG_EINVAL, "jpeg: file %d is %dx%d, the decoder reads %dx%d frames",
                i, hdr[i].width, hdr[i].height, d->width, d->height);
    JpegImage& im = d->h_img[i];
    memset(&im, 0, sizeof(im));
    im.ncomp = hdr[i].ncomp;
    im.hmax = hdr[i].hmax;
    im.vmax = hdr[i].vmax;
    long long blocks = 0;
    for (int c = 0; c < 3; ++c) {
      im.block_base[c] = blocks;
      if (c < hdr[i].ncomp) {
        im.h[c] = hdr[i].comp[c].h;
        im.v[c] = hdr[i].comp[c].v;
        im.bw[c] = hdr[i].bw[c];
        im.bh[c] = hdr[i].bh[c];
        im.coef_off[c] = (long long)coef_total;
        im.plane_off[c] = (long long)plane_total;
        const size_t nb = (size_t)im.bw[c] * im.bh[c];
        coef_total += nb * 64;
        plane_total += nb * 64;
        blocks += (long long)nb;
        memcpy(d->h_qt + ((size_t)i * 3 + c) * 64, hdr[i].qt[hdr[i].comp[c].tq], 128);
      } else {
        im.h[c] = im.v[c] = 1;
      }
    }
    im.block_base[3] = blocks;
    max_blocks = std::max(max_blocks, blocks);
  }
  SAG_REQUIRE(coef_total <= d->coef_cap * (size_t)d->max_frames, SAG_ENOMEM, "jpeg: coefficient staging too small");
  // (device entropy decoding leaves the host ~30 us of work per frame: a pool only pays when the caller asks for one)
  int nt = threads > 0 ? threads : (d->device_huffman ? 1 : (int)std::thread::hardware_concurrency());
  nt = std::max(1, std::min(std::min(nt, n), 32));
  auto run_pool = [&](const std::function<void(int)>& task) {      // task(i) for every frame, on nt threads
    std::atomic<int> next(0);
    auto work = [&]() { for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) task(i); };
    if (nt == 1) { work(); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < nt - 1; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
  };
  SAG_CHECK_CUDA(cudaMemcpyAsync(d->d_qt, d->h_qt, (size_t)n * 3 * 64 * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
  SAG_CHECK_CUDA(cudaMemcpyAsync(d->d_img, d->h_img, (size_t)n * sizeof(JpegImage), cudaMemcpyHostToDevice, st));
  if (d->device_huffman) {
    // host: unstuff the scans (memchr-paced copies) and lay out the subsequences; device: everything else
    size_t bound = 0;
    for (int i = 0; i < n; ++i) bound += (hdr[i].ecs_size + 16 + 15) / 16 * 16;
    SAG_TRY(grow(&d->h_stream, &d->d_stream, &d->stream_cap, bound));
    std::vector<size_t> off(n);
    size_t o = 0;
    for (int i = 0; i < n; ++i) { off[i] = o; o += (hdr[i].ecs_size + 16 + 15) / 16 * 16; }
    std::vector<std::vector<JpegSub>> subs(n);
    run_pool([&](int i) {
      std::vector<uint8_t> stream;
      std::vector<int> seg;
      unstuff(hdr[i], stream, seg);
      memcpy(d->h_stream + off[i], stream.data(), stream.size());
      memset(d->h_stream + off[i] + stream.size(), 0, (hdr[i].ecs_size + 16 + 15) / 16 * 16 - stream.size());
      plan_stream(hdr[i], seg, (int)stream.size(), &d->h_js[i], &subs[i]);
      d->h_js[i].data_off = (long long)off[i];
    });
    size_t n_subs = 0;
    for (int i = 0; i < n; ++i) { d->h_js[i].sub0 = (int)n_subs; n_subs += subs[i].size(); }
    if (n_subs > d->sub_cap) {
      if (d->d_scratch) SAG_CHECK_CUDA(cudaFree(d->d_scratch));
      d->d_scratch = nullptr;
    }
    SAG_TRY(grow(&d->h_subs, &d->d_subs, &d->sub_cap, n_subs));
    if (!d->d_scratch) SAG_CHECK_CUDA(cudaMalloc(&d->d_scratch, kScratchPerSub * d->sub_cap + 64));
    for (int i = 0; i < n; ++i) memcpy(d->h_subs + d->h_js[i].sub0, subs[i].data(), subs[i].size() * sizeof(JpegSub));
    // the batch's distinct Huffman tables (frames of one encoder share theirs)
    std::vector<const HuffTable*> uniq;
    auto row_of = [&](const HuffTable& t) {
      for (size_t k = 0; k < uniq.size(); ++k)
        if (memcmp(uniq[k], &t, sizeof(HuffTable)) == 0) return (int)k;
      uniq.push_back(&t);
      return (int)uniq.size() - 1;
    };
    for (int i = 0; i < n; ++i)
      for (int c = 0; c < 3; ++c) {
        const int cc = std::min(c, hdr[i].ncomp - 1);
        d->h_js[i].dc_tab[c] = row_of(hdr[i].dc[hdr[i].comp[cc].td]);
        d->h_js[i].ac_tab[c] = row_of(hdr[i].ac[hdr[i].comp[cc].ta]);
      }
    SAG_TRY(grow(&d->h_tables, &d->d_tables, &d->table_cap, uniq.size()));
    for (size_t k = 0; k < uniq.size(); ++k) memcpy(d->h_tables + k, uniq[k], sizeof(HuffTable));
    SAG_CHECK_CUDA(cudaMemcpyAsync(d->d_stream, d->h_stream, o, cudaMemcpyHostToDevice, st));
    SAG_CHECK_CUDA(cudaMemcpyAsync(d->d_subs, d->h_subs, n_subs * sizeof(JpegSub), cudaMemcpyHostToDevice, st));
    SAG_CHECK_CUDA(cudaMemcpyAsync(d->d_js, d->h_js, (size_t)n * sizeof(JpegStream), cudaMemcpyHostToDevice, st));
    SAG_CHECK_CUDA(cudaMemcpyAsync(d->d_tables, d->h_tables, uniq.size() * sizeof(HuffTable), cudaMemcpyHostToDevice, st));
    SAG_CHECK_CUDA(cudaEventRecord(d->staged, st));
    d->staged_pending = true;
    SAG_CHECK_CUDA(cudaMemsetAsync(d->d_coef, 0, coef_total * sizeof(int16_t), st));
    jpeg_huffman_kernel<<<n, kHuffThreads, 0, st>>>(d->d_js, d->d_subs, d->d_tables, d->d_img, d->d_stream, carve_scratch(d->d_scratch, d->sub_cap),
                                                    d->d_coef, d->d_rounds);
    SAG_LAUNCH_CHECK();
    d->last_n = n;
  } else {
    // entropy decoding on the host: one file per task
    std::atomic<int> failed(0);
    std::string first_error;
    std::atomic_flag err_lock = ATOMIC_FLAG_INIT;
    run_pool([&](int i) {
      const JpegImage& im = d->h_img[i];
      int16_t* cp[3] = {nullptr, nullptr, nullptr};
      size_t cnt = 0;
      for (int c = 0; c < im.ncomp; ++c) { cp[c] = d->h_coef + im.coef_off[c]; cnt += (size_t)im.bw[c] * im.bh[c] * 64; }
      memset(cp[0], 0, cnt * sizeof(int16_t));                  // (a frame's components are contiguous)
      if (decode_scan(hdr[i], cp) != SAG_OK) {
        failed.store(1);
        while (err_lock.test_and_set()) {}
        if (first_error.empty()) first_error = "file " + std::to_string(i) + ": " + sag_last_error();
        err_lock.clear();
      }
    });
    SAG_REQUIRE(!failed.load(), SAG_EINVAL, "%s", first_error.c_str());
    SAG_CHECK_CUDA(cudaMemcpyAsync(d->d_coef, d->h_coef, coef_total * sizeof(int16_t), cudaMemcpyHostToDevice, st));
    SAG_CHECK_CUDA(cudaEventRecord(d->staged, st));
    d->staged_pending = true;
    d->last_n = 0;
  }
  jpeg_idct_kernel<<<dim3((unsigned)((max_blocks + kIdctBlocksPerCta - 1) / kIdctBlocksPerCta), n), 256, 0, st>>>(d->d_img, d->d_coef, d->d_qt,
                                                                                                                 d->d_planes);
  SAG_LAUNCH_CHECK();
  jpeg_rgb_kernel<<<dim3((unsigned)((d->width + 255) / 256), d->height, n), 64, 0, st>>>(d->d_img, d->d_planes, d->width, d->height, frames);
  SAG_LAUNCH_CHECK();
  return SAG_OK;
}

}  // extern "C"
