// Geometry builders: map the reference's layer semantics (TF-1.4 conv / conv2d_transpose)
// onto the GatherGeom consumed by the contraction kernels.
#include "common.cuh"

namespace sag {

thread_local std::string g_err;
thread_local int g_launch_count = 0;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

const char* last_error_cstr() { return g_err.c_str(); }

thread_local Profiler* g_prof = nullptr;

bool pdl_enabled() {
  static const bool on = [] { const char* v = getenv("SAG_PDL"); return v == nullptr || atoi(v) != 0; }();
  return on;
}

void Profiler::clear() {
  for (auto& r : recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  recs.clear();
}

ProfScope::ProfScope(int cat, double flops, double bytes, cudaStream_t s, const char* name, double issued, int tile, int split) : st(s) {
  if (g_prof == nullptr || !g_prof->on) return;
  ProfRec r;
  r.cat = cat; r.flops = flops; r.bytes = bytes;
  r.issued = issued < 0 ? flops : issued;
  r.tile = tile; r.split = split;
  snprintf(r.name, sizeof(r.name), "%s", name ? name : "");
  if (cudaEventCreate(&r.e0) != cudaSuccess) return;
  if (cudaEventCreate(&r.e1) != cudaSuccess) { cudaEventDestroy(r.e0); return; }
  cudaEventRecord(r.e0, st);
  g_prof->recs.push_back(r);
  idx = (int)g_prof->recs.size() - 1;
}

ProfScope::~ProfScope() {
  if (idx >= 0 && g_prof != nullptr) cudaEventRecord(g_prof->recs[idx].e1, st);
}

// tf.nn.convolution (core.py:206): cross-correlation, VALID or TF-SAME (pad_before = total/2).
int make_conv_geom(GatherGeom* g, int n, int h, int w, int cin, int64_t x_ld, int kh, int kw, int cout, int sh, int sw,
                   int same_pad, int64_t y_ld, int* oh, int* ow) {
  SAG_REQUIRE(kh * kw <= kMaxTaps, SAG_EINVAL, "conv kernel %dx%d has more than %d taps", kh, kw, kMaxTaps);
  SAG_REQUIRE(n > 0 && h > 0 && w > 0 && cin > 0 && cout > 0 && sh > 0 && sw > 0, SAG_EINVAL, "conv: bad dims");
  int pt = 0, pl = 0, OH, OW;
  if (same_pad) {
    pt = same_pad_before(h, kh, sh, &OH);
    pl = same_pad_before(w, kw, sw, &OW);
  } else {
    SAG_REQUIRE(h >= kh && w >= kw, SAG_EINVAL, "conv VALID: input %dx%d smaller than kernel %dx%d", h, w, kh, kw);
    OH = (h - kh) / sh + 1;
    OW = (w - kw) / sw + 1;
  }
  memset(g, 0, sizeof(*g));
  g->N = n; g->H = h; g->W = w; g->Cin = cin; g->x_ld = x_ld;
  g->PH = OH; g->PW = OW; g->isy = sh; g->isx = sw;
  g->oy0 = 0; g->ox0 = 0; g->osy = 1; g->osx = 1;
  g->y_sc = 1; g->y_sw = y_ld; g->y_sh = (int64_t)OW * y_ld; g->y_sn = (int64_t)OH * OW * y_ld;
  g->Cout = cout;
  g->T = kh * kw;
  for (int r = 0; r < kh; ++r)
    for (int s = 0; s < kw; ++s) {
      int t = r * kw + s;
      g->dy[t] = (short)(r - pt);
      g->dx[t] = (short)(s - pl);
      g->widx[t] = (short)t;
    }
  *oh = OH; *ow = OW;
  return SAG_OK;
}

// tf.nn.conv2d_transpose VALID (core.py:139-140): y[o] += x[i] * w[p], o = i*s + p.  Output phase py: o = py + s*u,
// taps p = py + s*t (t < ceil((k-py)/s)), input i = u - t.  Rows restricted to full-output rows [row0,row1).
int make_deconv_phase_geom(GatherGeom* g, int n, int h, int w, int cin, int64_t x_ld, int kh, int kw, int cout, int sh,
                           int sw, int py, int px, int row0, int row1, int64_t y_sn, int64_t y_sh, int64_t y_sw,
                           int64_t y_sc) {
  const int OHf = (h - 1) * sh + kh, OWf = (w - 1) * sw + kw;
  if (row1 > OHf) row1 = OHf;
  int ty = (kh - py + sh - 1) / sh, tx = (kw - px + sw - 1) / sw;      // taps of this phase
  // stride > kernel: phases no tap reaches still exist in the output and receive the bias only (T == 0)
  if (ty <= 0 || tx <= 0) { ty = 0; tx = 0; }
  SAG_REQUIRE(ty * tx <= kMaxTaps, SAG_EINVAL, "deconv phase has too many taps");
  // rows o = py + sh*u in [row0,row1)
  int u0 = (row0 - py + sh - 1) / sh;
  if (row0 - py < 0) u0 = 0;
  int u1 = (row1 - 1 - py) / sh;           // inclusive
  if (row1 - 1 - py < 0) return 1;
  int PH = u1 - u0 + 1;
  int PW = (OWf - 1 - px) / sw + 1;
  if (OWf - 1 - px < 0) return 1;
  if (PH <= 0 || PW <= 0) return 1;
  memset(g, 0, sizeof(*g));
  g->N = n; g->H = h; g->W = w; g->Cin = cin; g->x_ld = x_ld;
  g->PH = PH; g->PW = PW; g->isy = 1; g->isx = 1;
  g->oy0 = py + sh * u0 - row0; g->ox0 = px; g->osy = sh; g->osx = sw;
  g->y_sn = y_sn; g->y_sh = y_sh; g->y_sw = y_sw; g->y_sc = y_sc;
  g->Cout = cout;
  g->T = ty * tx;
  for (int a = 0; a < ty; ++a)
    for (int b = 0; b < tx; ++b) {
      int t = a * tx + b;
      g->dy[t] = (short)(u0 - a);
      g->dx[t] = (short)(-b);
      g->widx[t] = (short)((py + sh * a) * kw + (px + sw * b));
    }
  return SAG_OK;
}

}  // namespace sag
