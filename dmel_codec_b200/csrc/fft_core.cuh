// Register-resident FFT building blocks for the fused dMel kernel.
//
// A warp computes one 1024-point complex FFT as 32 x 32 (Cooley-Tukey, two
// radix-32 passes).  Every lane keeps 32 complex points in registers, does a
// fully unrolled radix-32 DFT on them, and the single 32x32 transpose between
// the two passes goes through a padded shared-memory tile (one STS.64 sweep,
// one LDS.128 sweep, both bank-conflict free).  Real input rides on that core
// either as two frames packed into one complex signal (n_fft = 1024) or as one
// frame folded to half length (n_fft = 2048); see logmel_kernel.cuh.
//
// Replaces, together with logmel_kernel.cuh, the torch.stft call at reference
// dmel_codec/utils/spectrogram.py:64-75.
//
// Everything here is __host__ __device__ so tests/host_emul.cu can run the same
// code on the CPU, lane by lane, against a float64 DFT.
#pragma once
#include <cuda_runtime.h>
#include <utility>

namespace dmel {

#define DMEL_HD __host__ __device__ __forceinline__

// bit reversal of a 5-bit index
DMEL_HD constexpr int brev5(int x) {
  return ((x & 1) << 4) | ((x & 2) << 2) | (x & 4) | ((x & 8) >> 2) | ((x & 16) >> 4);
}

// cos(2*pi*q/32) for q = 0..8, correctly rounded to float
DMEL_HD constexpr float cos32_q(int q) {
  switch (q) {
    case 0: return 1.0f;
    case 1: return 0.98078528040323043f;
    case 2: return 0.92387953251128674f;
    case 3: return 0.83146961230254524f;
    case 4: return 0.70710678118654752f;
    case 5: return 0.55557023301960218f;
    case 6: return 0.38268343236508977f;
    case 7: return 0.19509032201612825f;
    default: return 0.0f;
  }
}
// W_32^q = cos32(q) - i * sin32(q), q = 0..15
DMEL_HD constexpr float cos32(int q) { return q <= 8 ? cos32_q(q) : -cos32_q(16 - q); }
DMEL_HD constexpr float sin32(int q) { return q <= 8 ? cos32_q(8 - q) : cos32_q(q - 8); }

DMEL_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
DMEL_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
DMEL_HD float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(-a.y, b.y, a.x * b.x), fmaf(a.y, b.x, a.x * b.y));
}

// d * W_32^Q with the trivial rotations folded away at compile time
template <int Q>
DMEL_HD float2 mul_w32(float2 d) {
  constexpr float kR2 = 0.70710678118654752f;
  if constexpr (Q == 0) {
    return d;
  } else if constexpr (Q == 8) {  // -i
    return make_float2(d.y, -d.x);
  } else if constexpr (Q == 4) {  // (1 - i)/sqrt2
    return make_float2(kR2 * (d.x + d.y), kR2 * (d.y - d.x));
  } else if constexpr (Q == 12) {  // (-1 - i)/sqrt2
    return make_float2(kR2 * (d.y - d.x), -kR2 * (d.x + d.y));
  } else {
    constexpr float c = cos32(Q), s = sin32(Q);
    return make_float2(fmaf(d.y, s, d.x * c), fmaf(-d.x, s, d.y * c));
  }
}

// one decimation-in-frequency butterfly of span H, butterfly number I (0..15)
template <int H, int I>
DMEL_HD void dif_butterfly(float2 (&a)[32]) {
  constexpr int blk = I / H, j = I % H;
  constexpr int p = blk * 2 * H + j, q = p + H;
  const float2 u = a[p], w = a[q];
  a[p] = cadd(u, w);
  a[q] = mul_w32<j * (16 / H)>(csub(u, w));
}
template <int H, int... I>
DMEL_HD void dif_stage(float2 (&a)[32], std::integer_sequence<int, I...>) {
  (dif_butterfly<H, I>(a), ...);
}

// In-place forward 32-point DFT (kernel e^{-2 pi i nk/32}).  Result is left in
// bit-reversed order: X[k] == a[brev5(k)].
DMEL_HD void radix32(float2 (&a)[32]) {
  using seq = std::make_integer_sequence<int, 16>;
  dif_stage<16>(a, seq{});
  dif_stage<8>(a, seq{});
  dif_stage<4>(a, seq{});
  dif_stage<2>(a, seq{});
  dif_stage<1>(a, seq{});
}

// Shared-memory transpose tile of one warp: 32 rows of 32 complex, row pitch
// 34 complex = 272 B.  272/16 is odd, so eight lanes reading 16 B each at
// consecutive rows cover all 32 banks (LDS.128 conflict free); a row written
// by 32 lanes as 8-byte words is contiguous (STS.64 conflict free).
constexpr int kTilePitch = 34;                     // in float2
constexpr int kTileFloat2 = 32 * kTilePitch;       // 1088 float2 = 8704 B

// Pass 1 of the 1024-point FFT for lane n2:  v[n1] = z[32*n1 + n2] on entry.
// Leaves  Y[n2][k1] * W_1024^{n2*k1}  at tile[k1][n2].  tw[k1] = W_1024^{n2*k1}.
DMEL_HD void fft1024_pass1(float2 (&v)[32], const float2 (&tw)[32], float2* tile, int lane) {
  radix32(v);
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) {
    const float2 y = v[brev5(k1)];
    tile[k1 * kTilePitch + lane] = (k1 == 0) ? y : cmul(y, tw[k1]);
  }
}

// Pass 2 for lane k1: reads row k1 of the tile, leaves Z[k1 + 32*k2] in
// v[brev5(k2)].
DMEL_HD void fft1024_pass2(float2 (&v)[32], const float2* tile, int lane) {
  const float4* row = reinterpret_cast<const float4*>(tile + lane * kTilePitch);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float4 t = row[j];
    v[2 * j] = make_float2(t.x, t.y);
    v[2 * j + 1] = make_float2(t.z, t.w);
  }
  radix32(v);
}

// Which register of the lane that owns bin (1024 - k) holds it, for
// k = lane + 32*k2.  Lane 0 pairs with itself one slot later than the others.
DMEL_HD constexpr int mirror_slot(int k2, bool sender_is_lane0) {
  return sender_is_lane0 ? ((32 - k2) & 31) : (31 - k2);
}

DMEL_HD float fast_sqrt(float x) {
#ifdef __CUDA_ARCH__
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return sqrtf(x);
#endif
}

constexpr float kMagEps = 1e-9f;  // reference utils/spectrogram.py:76

// Two real frames packed as z = a + i*b.  A = Z[k], Bm = Z[N-k].
//   Xa[k] = (A + conj(Bm))/2 ,  Xb[k] = (A - conj(Bm))/(2i)
// Returns sqrt(|X|^2 + 1e-9) for both frames.
DMEL_HD void packed_pair_magnitudes(float2 A, float2 Bm, float& mag_a, float& mag_b) {
  const float sr = A.x + Bm.x, di = A.y - Bm.y;
  const float dr = A.x - Bm.x, si = A.y + Bm.y;
  mag_a = fast_sqrt(fmaf(0.25f, fmaf(sr, sr, di * di), kMagEps));
  mag_b = fast_sqrt(fmaf(0.25f, fmaf(dr, dr, si * si), kMagEps));
}

// One real frame of 2048 folded to z[n] = x[2n] + i*x[2n+1], Z = FFT_1024(z).
// With E = (A + conj(Bm))/2, O = (A - conj(Bm))/(2i), w = W_2048^k:
//   X[k] = E + w*O ,  X[1024-k] = conj(E - w*O)
DMEL_HD void folded_magnitudes(float2 A, float2 Bm, float2 w, float& mag_k, float& mag_mirror) {
  const float2 E = make_float2(0.5f * (A.x + Bm.x), 0.5f * (A.y - Bm.y));
  const float2 O = make_float2(0.5f * (A.y + Bm.y), -0.5f * (A.x - Bm.x));
  const float2 wo = cmul(w, O);
  const float2 p = cadd(E, wo), m = csub(E, wo);
  mag_k = fast_sqrt(fmaf(p.x, p.x, fmaf(p.y, p.y, kMagEps)));
  mag_mirror = fast_sqrt(fmaf(m.x, m.x, fmaf(m.y, m.y, kMagEps)));
}

}  // namespace dmel
