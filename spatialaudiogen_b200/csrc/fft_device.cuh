// Device-side shared-memory Stockham FFT (radices 2,3,4,5) shared by fft.cu and metrics.cu.
#pragma once
#include "common.cuh"

namespace sag {

constexpr int kMaxPasses = 8;
struct FftPlan {
  int n;
  int npass;
  unsigned radix_code;  // radix of pass s in bits [4s, 4s+4): register/constant-bank friendly (an indexed array would be
                        // copied to local memory inside the kernels)
  const float2* tw;     // exp(-2*pi*i*m/n), m in [0,n)
  const float* hann;    // periodic Hann of length n
};

int get_plan(int n, FftPlan* out);   // fft.cu: cached per (device, n); creates twiddle/Hann tables on first use

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// One Stockham pass of radix R over the block: in/out are shared-memory arrays of n complex values.
template <int R>
__device__ __forceinline__ void stockham_pass(const float2* __restrict__ in, float2* __restrict__ out, int n, int ns,
                                              const float2* __restrict__ tw) {
  const int nr = n / R;
  const int tstep = n / (ns * R);
  for (int j = threadIdx.x; j < nr; j += blockDim.x) {
    const int k = j % ns;
    float2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      v[r] = in[j + r * nr];
      if (r > 0 && ns > 1) v[r] = cmul(v[r], __ldg(tw + r * k * tstep));
    }
    if (R == 2) {
      float2 a = v[0], b = v[1];
      v[0] = cadd(a, b);
      v[1] = csub(a, b);
    } else if (R == 4) {
      float2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
      float2 c = cadd(v[1], v[3]), d = csub(v[1], v[3]);
      float2 dj = make_float2(d.y, -d.x);                 // -i * d
      v[0] = cadd(a, c);
      v[1] = cadd(b, dj);
      v[2] = csub(a, c);
      v[3] = csub(b, dj);
    } else if (R == 3) {
      const float s = 0.86602540378443864676f;            // sin(2pi/3)
      float2 t1 = cadd(v[1], v[2]);
      float2 t2 = make_float2(v[0].x - 0.5f * t1.x, v[0].y - 0.5f * t1.y);
      float2 d = csub(v[1], v[2]);
      float2 t3 = make_float2(s * d.y, -s * d.x);         // -i*s*d
      v[0] = cadd(v[0], t1);
      v[1] = cadd(t2, t3);
      v[2] = csub(t2, t3);
    } else if (R == 5) {
      const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;   // cos(2pi/5), cos(4pi/5)
      const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;    // sin(2pi/5), sin(4pi/5)
      float2 a1 = cadd(v[1], v[4]), b1 = csub(v[1], v[4]);
      float2 a2 = cadd(v[2], v[3]), b2 = csub(v[2], v[3]);
      float2 m1 = make_float2(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
      float2 m2 = make_float2(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
      float2 q1 = make_float2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y);
      float2 q2 = make_float2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y);
      float2 q1j = make_float2(q1.y, -q1.x), q2j = make_float2(q2.y, -q2.x);     // -i*q
      v[0] = make_float2(v[0].x + a1.x + a2.x, v[0].y + a1.y + a2.y);
      v[1] = cadd(m1, q1j);
      v[4] = csub(m1, q1j);
      v[2] = cadd(m2, q2j);
      v[3] = csub(m2, q2j);
    }
    const int base = (j / ns) * ns * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) out[base + r * ns] = v[r];
  }
}

// One Stockham pass of radix R over nf independent transforms stored back to back (frame g at offset g*n): the
// twiddles are fetched once per butterfly position and reused for every frame -- nf times fewer block-wide syncs per
// transform than running them one after the other.
template <int R>
__device__ __forceinline__ void stockham_pass_nf(const float2* __restrict__ in, float2* __restrict__ out, int n, int ns,
                                                 const float2* __restrict__ tw, int nf) {
  const int nr = n / R;
  const int tstep = n / (ns * R);
  for (int j = threadIdx.x; j < nr; j += blockDim.x) {
    const int k = j % ns;
    float2 w[R];
#pragma unroll
    for (int r = 1; r < R; ++r) w[r] = ns > 1 ? __ldg(tw + r * k * tstep) : make_float2(1.f, 0.f);
    const int base = (j / ns) * ns * R + k;
    for (int g = 0; g < nf; ++g) {
      const float2* ig = in + g * n;
      float2* og = out + g * n;
      float2 v[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        v[r] = ig[j + r * nr];
        if (r > 0 && ns > 1) v[r] = cmul(v[r], w[r]);
      }
      if (R == 2) {
        float2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
      } else if (R == 4) {
        float2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
        float2 c = cadd(v[1], v[3]), d = csub(v[1], v[3]);
        float2 dj = make_float2(d.y, -d.x);
        v[0] = cadd(a, c);
        v[1] = cadd(b, dj);
        v[2] = csub(a, c);
        v[3] = csub(b, dj);
      } else if (R == 3) {
        const float s = 0.86602540378443864676f;
        float2 t1 = cadd(v[1], v[2]);
        float2 t2 = make_float2(v[0].x - 0.5f * t1.x, v[0].y - 0.5f * t1.y);
        float2 d = csub(v[1], v[2]);
        float2 t3 = make_float2(s * d.y, -s * d.x);
        v[0] = cadd(v[0], t1);
        v[1] = cadd(t2, t3);
        v[2] = csub(t2, t3);
      } else {
        const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
        const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
        float2 a1 = cadd(v[1], v[4]), b1 = csub(v[1], v[4]);
        float2 a2 = cadd(v[2], v[3]), b2 = csub(v[2], v[3]);
        float2 m1 = make_float2(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
        float2 m2 = make_float2(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
        float2 q1 = make_float2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y);
        float2 q2 = make_float2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y);
        float2 q1j = make_float2(q1.y, -q1.x), q2j = make_float2(q2.y, -q2.x);
        v[0] = make_float2(v[0].x + a1.x + a2.x, v[0].y + a1.y + a2.y);
        v[1] = cadd(m1, q1j);
        v[4] = csub(m1, q1j);
        v[2] = cadd(m2, q2j);
        v[3] = csub(m2, q2j);
      }
#pragma unroll
      for (int r = 0; r < R; ++r) og[base + r * ns] = v[r];
    }
  }
}

// Forward FFT of nf transforms of n complex values each (frame g at buf0 + g*n). Returns the buffer holding the results.
__device__ __forceinline__ float2* block_fft_nf(float2* buf0, float2* buf1, const FftPlan& p, int nf) {
  int ns = 1;
  float2* in = buf0;
  float2* out = buf1;
  for (int s = 0; s < p.npass; ++s) {
    __syncthreads();
    const int R = (int)((p.radix_code >> (4 * s)) & 15u);
    if (R == 4) stockham_pass_nf<4>(in, out, p.n, ns, p.tw, nf);
    else if (R == 2) stockham_pass_nf<2>(in, out, p.n, ns, p.tw, nf);
    else if (R == 3) stockham_pass_nf<3>(in, out, p.n, ns, p.tw, nf);
    else stockham_pass_nf<5>(in, out, p.n, ns, p.tw, nf);
    ns *= R;
    float2* t = in; in = out; out = t;
  }
  __syncthreads();
  return in;
}

// Forward FFT of the n complex values in buf0 (shared). Returns the buffer holding the result.
__device__ __forceinline__ float2* block_fft(float2* buf0, float2* buf1, const FftPlan& p) {
  int ns = 1;
  float2* in = buf0;
  float2* out = buf1;
  for (int s = 0; s < p.npass; ++s) {
    __syncthreads();
    const int R = (int)((p.radix_code >> (4 * s)) & 15u);
    if (R == 4) stockham_pass<4>(in, out, p.n, ns, p.tw);
    else if (R == 2) stockham_pass<2>(in, out, p.n, ns, p.tw);
    else if (R == 3) stockham_pass<3>(in, out, p.n, ns, p.tw);
    else stockham_pass<5>(in, out, p.n, ns, p.tw);
    ns *= R;
    float2* t = in; in = out; out = t;
  }
  __syncthreads();
  return in;
}

}  // namespace sag
