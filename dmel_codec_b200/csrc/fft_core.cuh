// Register-resident FFT building blocks for the fused dMel kernel.
//
// A real frame of n_fft samples is folded to n_fft/2 complex points
// z[n] = x[2n] + i*x[2n+1]; one warp computes Z = FFT(z) in two register passes
// around a single shared-memory transpose, then unfolds Z into the n_fft/2+1
// magnitudes.  Every frame has its own FFT, so a quiet frame never inherits
// rounding noise from a loud neighbour.
//
//   n_fft = 1024 : 512 = 16 x 32.  Pass 1: lane n2 does a radix-16 DFT over
//                  n1 (16 points in registers).  Pass 2: the 32-point DFT of
//                  row k1 is shared by the lane pair (k1, k1+16): each lane a
//                  radix-16 over one parity of n2, then one cross-lane radix-2.
//   n_fft = 2048 : 1024 = 32 x 32, radix-32 in registers in both passes.
//
// Replaces, together with logmel_kernel.cuh, the torch.stft call at reference
// dmel_codec/utils/spectrogram.py:64-75 and the magnitude at :76.
//
// The arithmetic is __host__ __device__ and split at every cross-lane exchange,
// so tests/host_emul.cu runs the same code on the CPU lane by lane against a
// float64 DFT.
#pragma once
#include <cuda_runtime.h>
#include <utility>

namespace dmel {

#define DMEL_HD __host__ __device__ __forceinline__

DMEL_HD constexpr int brev5(int x) {
  return ((x & 1) << 4) | ((x & 2) << 2) | (x & 4) | ((x & 8) >> 2) | ((x & 16) >> 4);
}
DMEL_HD constexpr int brev4(int x) {
  return ((x & 1) << 3) | ((x & 2) << 1) | ((x & 4) >> 1) | ((x & 8) >> 3);
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

// ---- complex arithmetic on (re, im) pairs ---------------------------------------
// sm_100 can execute add/mul/fma on an aligned register PAIR in one instruction (PTX
// add/mul/fma.rn.f32x2, SASS FADD2/FMUL2/FFMA2) with free operand swizzles (swap halves, negate
// one half, broadcast a scalar): a complex add is one instruction, a complex multiply two.
// A packed op holds the FMA pipe for two cycles (benchmarks/fp_issue_rate.cu: same flops per cycle
// as scalar) but takes ONE issue slot, and the slot it frees is usable by the other pipes
// (benchmarks/f32x2_coissue.cu: 4 FFMA2 + 16 ALU instructions issue in 24 cycles, 8 FFMA + 16 ALU in
// 34).  The fused kernel is bound by issue slots, so packed arithmetic is the default (-4.4 % time,
// profiles/history.md); -DDMEL_SCALAR_F32 builds the scalar form.  Both forms round identically
// (each half is an IEEE fma), and the host versions spell out the same operations in the same order
// for tests/host_emul.cu.
#if defined(__CUDA_ARCH__) && !defined(DMEL_SCALAR_F32)
#define DMEL_PACKED_F32X2 1
#endif
#ifdef DMEL_PACKED_F32X2
__device__ __forceinline__ unsigned long long f2_pack(float2 a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ float2 f2_unpack(unsigned long long r) {
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(r));
  return d;
}
#endif
DMEL_HD float2 f2_add(float2 a, float2 b) {
#ifdef DMEL_PACKED_F32X2
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(d);
#else
  return make_float2(a.x + b.x, a.y + b.y);
#endif
}
DMEL_HD float2 f2_mul(float2 a, float2 b) {
#ifdef DMEL_PACKED_F32X2
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(d);
#else
  return make_float2(a.x * b.x, a.y * b.y);
#endif
}
DMEL_HD float2 f2_fma(float2 a, float2 b, float2 c) {
#ifdef DMEL_PACKED_F32X2
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)), "l"(f2_pack(c)));
  return f2_unpack(d);
#else
  return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
DMEL_HD float2 f2_swap(float2 a) { return make_float2(a.y, a.x); }

DMEL_HD float2 cadd(float2 a, float2 b) { return f2_add(a, b); }
DMEL_HD float2 csub(float2 a, float2 b) { return f2_add(a, make_float2(-b.x, -b.y)); }
// a * b = a*(b.x, b.x) + swap(a)*(-b.y, b.y)
DMEL_HD float2 cmul(float2 a, float2 b) {
  return f2_fma(f2_swap(a), make_float2(-b.y, b.y), f2_mul(a, make_float2(b.x, b.x)));
}
// d * (c - i s) = d*(c, c) + swap(d)*(s, -s)
DMEL_HD float2 cmul_conj_cs(float2 d, float c, float s) {
  return f2_fma(f2_swap(d), make_float2(s, -s), f2_mul(d, make_float2(c, c)));
}

// p + q * (c - i s): the twiddle multiply folded into the butterfly's add (two packed / four scalar FMAs)
DMEL_HD float2 cfma_conj_cs(float2 p, float2 q, float c, float s) {
  return f2_fma(f2_swap(q), make_float2(s, -s), f2_fma(q, make_float2(c, c), p));
}
// p + q * w
DMEL_HD float2 cfma(float2 p, float2 q, float2 w) {
  return f2_fma(f2_swap(q), make_float2(-w.y, w.y), f2_fma(q, make_float2(w.x, w.x), p));
}
// 2p - r: the second output of a butterfly whose first output r = p + t is known (p - t = 2p - r)
DMEL_HD float2 twice_minus(float2 p, float2 r) { return f2_fma(p, make_float2(2.f, 2.f), make_float2(-r.x, -r.y)); }

// d * W_32^Q with the trivial rotations folded away at compile time
template <int Q>
DMEL_HD float2 mul_w32(float2 d) {
  if constexpr (Q == 0) {
    return d;
  } else if constexpr (Q == 8) {  // -i : (d.y, -d.x); swap and sign ride on the consumer's operand modifiers
    return make_float2(d.y, -d.x);
  } else {
    return cmul_conj_cs(d, cos32(Q), sin32(Q));
  }
}

// ---- radix-32 / radix-16 decimation-in-frequency butterflies ------------------
template <int N, int H, int I>
DMEL_HD void dif_butterfly(float2 (&a)[N]) {
  constexpr int blk = I / H, j = I % H;
  constexpr int p = blk * 2 * H + j, q = p + H;
  const float2 u = a[p], w = a[q];
  a[p] = cadd(u, w);
  a[q] = mul_w32<j * (16 / H)>(csub(u, w));  // W_{2H}^j == W_32^{j*16/H}
}
template <int N, int H, int... I>
DMEL_HD void dif_stage(float2 (&a)[N], std::integer_sequence<int, I...>) {
  (dif_butterfly<N, H, I>(a), ...);
}

// In-place forward 32-point DFT (kernel e^{-2 pi i nk/32}); X[k] == a[brev5(k)].
DMEL_HD void radix32(float2 (&a)[32]) {
  using seq = std::make_integer_sequence<int, 16>;
  dif_stage<32, 16>(a, seq{});
  dif_stage<32, 8>(a, seq{});
  dif_stage<32, 4>(a, seq{});
  dif_stage<32, 2>(a, seq{});
  dif_stage<32, 1>(a, seq{});
}
// Decimation-in-time butterfly with the twiddle folded into the add:
//   a[p] <- a[p] + W a[q],   a[q] <- a[p] - W a[q] = 2 a[p] - (a[p] + W a[q]),   W = W_{2H}^j = W_32^{j*16/H}
// six FMAs where multiply-then-add/subtract takes eight; trivial W stay four adds.
template <int N, int H, int I>
DMEL_HD void dit_butterfly(float2 (&a)[N]) {
  constexpr int blk = I / H, j = I % H;
  constexpr int p = blk * 2 * H + j, q = p + H;
  constexpr int Q = j * (16 / H);
  const float2 u = a[p], w = a[q];
  if constexpr (Q == 0) {
    a[p] = cadd(u, w);
    a[q] = csub(u, w);
  } else if constexpr (Q == 8) {  // W = -i: W w = (w.y, -w.x)
    a[p] = make_float2(u.x + w.y, u.y - w.x);
    a[q] = make_float2(u.x - w.y, u.y + w.x);
  } else {
    const float2 r = cfma_conj_cs(u, w, cos32(Q), sin32(Q));
    a[p] = r;
    a[q] = twice_minus(u, r);
  }
}
template <int N, int H, int... I>
DMEL_HD void dit_stage(float2 (&a)[N], std::integer_sequence<int, I...>) {
  (dit_butterfly<N, H, I>(a), ...);
}
// In-place forward 16-point DFT; X[k] == a[brev4(k)].  Decimation in time on a bit-reversed
// copy; every index is a compile-time constant, so both permutations are register renamings.
DMEL_HD void radix16(float2 (&a)[16]) {
  using seq = std::make_integer_sequence<int, 8>;
  float2 b[16];
#pragma unroll
  for (int n = 0; n < 16; ++n) b[brev4(n)] = a[n];
  dit_stage<16, 1>(b, seq{});
  dit_stage<16, 2>(b, seq{});
  dit_stage<16, 4>(b, seq{});
  dit_stage<16, 8>(b, seq{});
#pragma unroll
  for (int k = 0; k < 16; ++k) a[brev4(k)] = b[k];
}

// Shared-memory transpose tile of one warp: rows of 32 complex, row pitch
// 34 complex = 272 B.  272/16 is odd, so eight lanes reading 16 B each at
// consecutive rows cover all 32 banks (LDS.128 conflict free); a row written
// by 32 lanes as 8-byte words is contiguous (STS.64 conflict free).
constexpr int kTilePitch = 34;                  // in float2
constexpr int kTile512 = 16 * kTilePitch;       // 544 float2 = 4352 B  (n_fft 1024)
constexpr int kTile1024 = 32 * kTilePitch;      // 1088 float2 = 8704 B (n_fft 2048)

// natural log via MUFU.LG2 (inputs here are clamped to >= 1e-5, so no denormal handling)
DMEL_HD float fast_log(float x) {
#ifdef __CUDA_ARCH__
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r * 0.69314718055994531f;
#else
  return logf(x);
#endif
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

// Unfold one bin pair of a real frame of 2C samples from Z = FFT_C(z):
// with A = Z[k], Bm = Z[C-k], E = (A + conj Bm)/2, O = (A - conj Bm)/(2i),
// w = W_{2C}^k:   X[k] = E + w*O ,  X[C-k] = conj(E - w*O).
// Returns sqrt(|X|^2 + 1e-9) for both bins (the 1/2 is folded into the square).
DMEL_HD void folded_magnitudes(float2 A, float2 Bm, float2 w, float& mag_k, float& mag_mirror) {
  const float2 P = f2_fma(Bm, make_float2(1.f, -1.f), A);                            // 2E = (A.x+B.x, A.y-B.y)
  const float2 Q = f2_fma(f2_swap(A), make_float2(1.f, -1.f), f2_swap(Bm));          // 2O = (A.y+B.y, B.x-A.x)
  const float2 p = cfma(P, Q, w), m = twice_minus(P, p);
  const float2 pp = f2_mul(p, p), mm = f2_mul(m, m);
  mag_k = fast_sqrt(fmaf(0.25f, pp.x + pp.y, kMagEps));
  mag_mirror = fast_sqrt(fmaf(0.25f, mm.x + mm.y, kMagEps));
}

// =============================================================================
// n_fft = 1024 : 512-point complex FFT, 16 points per lane
// =============================================================================
// Pass 1, lane n2:  v[n1] = z[32*n1 + n2] on entry.  Leaves
// Y[n2][k1] * W_512^{n2*k1} at tile[k1][col(n2)], where even n2 fill columns
// 0..15 and odd n2 columns 16..31 so pass 2 reads one parity contiguously.
// The odd block is rotated by 8 columns: a 64-bit shared store is served per
// half-warp, and without the rotation the 8 even and 8 odd lanes of a half-warp
// would land on the same 16 banks.  Pass 2 therefore sees the odd-parity inputs
// circularly shifted by 8, i.e. its outputs are (-1)^q G_1[q]; combine_one folds
// that sign into its constants.  tw[k1] = W_512^{n2*k1}.
DMEL_HD void fft512_pass1(float2 (&v)[16], const float2 (&tw)[16], float2* tile, int lane) {
  radix16(v);
  const int odd = lane & 1;
  const int col = odd * 16 + (((lane >> 1) + 8 * odd) & 15);
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) {
    const float2 y = v[brev4(k1)];
    tile[k1 * kTilePitch + col] = (k1 == 0) ? y : cmul(y, tw[k1]);
  }
}

// Same pass with the 15 twiddles rebuilt per frame from four of them (w1, w2, w4, w8 =
// W_512^{n2*{1,2,4,8}}): 11 extra complex multiplies buy back 22 registers, which is what lets a
// third CTA fit on the SM.
DMEL_HD void fft512_pass1_pow(float2 (&v)[16], float2 w1, float2 w2, float2 w4, float2 w8, float2* tile, int lane) {
  radix16(v);
  const int odd = lane & 1;
  float2* col = tile + odd * 16 + (((lane >> 1) + 8 * odd) & 15);
  const float2 w3 = cmul(w1, w2);
  col[0 * kTilePitch] = v[brev4(0)];
  col[8 * kTilePitch] = cmul(v[brev4(8)], w8);
  col[1 * kTilePitch] = cmul(v[brev4(1)], w1);
  col[9 * kTilePitch] = cmul(cmul(v[brev4(9)], w1), w8);
  col[2 * kTilePitch] = cmul(v[brev4(2)], w2);
  col[10 * kTilePitch] = cmul(cmul(v[brev4(10)], w2), w8);
  col[3 * kTilePitch] = cmul(v[brev4(3)], w3);
  col[11 * kTilePitch] = cmul(cmul(v[brev4(11)], w3), w8);
  col[4 * kTilePitch] = cmul(v[brev4(4)], w4);
  col[12 * kTilePitch] = cmul(cmul(v[brev4(12)], w4), w8);
  const float2 w5 = cmul(w4, w1), w6 = cmul(w4, w2), w7 = cmul(w4, w3);
  col[5 * kTilePitch] = cmul(v[brev4(5)], w5);
  col[13 * kTilePitch] = cmul(cmul(v[brev4(13)], w5), w8);
  col[6 * kTilePitch] = cmul(v[brev4(6)], w6);
  col[14 * kTilePitch] = cmul(cmul(v[brev4(14)], w6), w8);
  col[7 * kTilePitch] = cmul(v[brev4(7)], w7);
  col[15 * kTilePitch] = cmul(cmul(v[brev4(15)], w7), w8);
}

// Pass 2, lane = k1 + 16*h: radix-16 over the n2 of parity h of row k1.
// Leaves G_0[q] (h = 0) or (-1)^q G_1[q] (h = 1) in v[brev4(q)], where
// G_h[q] = sum_m Y'[2m+h] W_16^{mq}.
DMEL_HD void fft512_pass2(float2 (&v)[16], const float2* tile, int lane) {
  const float4* row = reinterpret_cast<const float4*>(tile + (lane & 15) * kTilePitch + (lane >> 4) * 16);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 t = row[j];
    v[2 * j] = make_float2(t.x, t.y);
    v[2 * j + 1] = make_float2(t.z, t.w);
  }
  radix16(v);
}

// Cross-lane radix-2 that finishes the 32-point DFT of a row:
//   Z'[q + 16 r] = G_0[q] + (-1)^r W_32^q G_1[q].
// Lane h keeps the q of parity h (q = 2j + h) and ships the other parity to its
// partner (lane ^ 16).  combine_send picks what to ship, combine_finish consumes
// what arrived:  zlo[j] = Z'[2j+h],  zhi[j] = Z'[2j+h+16],  i.e. bins
// k = lane + 32*j  and  k + 256.
DMEL_HD void combine_send(const float2 (&v)[16], int h, float2 (&send)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) send[j] = h ? v[brev4(2 * j)] : v[brev4(2 * j + 1)];
}
template <int J>
DMEL_HD void combine_one(const float2 (&v)[16], const float2 (&recv)[8], int h, float2 (&zlo)[8], float2 (&zhi)[8]) {
  const float2 a = h ? recv[J] : v[brev4(2 * J)];        // G_0[q]
  const float2 b = h ? v[brev4(2 * J + 1)] : recv[J];    // G_1[q] for h = 0 (q even), -G_1[q] for h = 1 (q odd)
  const float c = h ? -cos32(2 * J + 1) : cos32(2 * J);  // W_32^q = c - i s, sign of the rotated read folded in
  const float s = h ? -sin32(2 * J + 1) : sin32(2 * J);
  zlo[J] = cfma_conj_cs(a, b, c, s);
  zhi[J] = twice_minus(a, zlo[J]);
}
template <int... J>
DMEL_HD void combine_all(const float2 (&v)[16], const float2 (&recv)[8], int h, float2 (&zlo)[8], float2 (&zhi)[8],
                         std::integer_sequence<int, J...>) {
  (combine_one<J>(v, recv, h, zlo, zhi), ...);
}
DMEL_HD void combine_finish(const float2 (&v)[16], const float2 (&recv)[8], int h, float2 (&zlo)[8], float2 (&zhi)[8]) {
  combine_all(v, recv, h, zlo, zhi, std::make_integer_sequence<int, 8>{});
}

// Unfold needs Z[512-k] for k = lane + 32*j, j = 0..7.  It lives in lane
// mirror_lane512(lane) as zhi[7-j]; lane 0 pairs with itself one slot later
// (and with its own zlo[0] for k = 0).
DMEL_HD int mirror_lane512(int lane) { return ((16 - (lane & 15)) + 16 * (1 - (lane >> 4))) & 31; }
DMEL_HD void mirror_send512(const float2 (&zlo)[8], const float2 (&zhi)[8], int lane, float2 (&send)[8]) {
  send[0] = (lane == 0) ? zlo[0] : zhi[7];
#pragma unroll
  for (int j = 1; j < 8; ++j) send[j] = (lane == 0) ? zhi[8 - j] : zhi[7 - j];
}
// W_1024^{lane + 32 j} = base * W_32^j with base = W_1024^lane
template <int J>
DMEL_HD void unfold_one512(const float2 (&zlo)[8], const float2 (&recv)[8], float2 base, float* mrow, int lane) {
  float mk, mm;
  folded_magnitudes(zlo[J], recv[J], mul_w32<J>(base), mk, mm);
  mrow[lane + 32 * J] = mk;
  mrow[512 - lane - 32 * J] = mm;
}
template <int... J>
DMEL_HD void unfold_all512(const float2 (&zlo)[8], const float2 (&recv)[8], float2 base, float* mrow, int lane,
                           std::integer_sequence<int, J...>) {
  (unfold_one512<J>(zlo, recv, base, mrow, lane), ...);
}
// Writes the 513 magnitudes of the frame into mrow[0..512] (this lane's share).
DMEL_HD void unfold_store512(const float2 (&zlo)[8], const float2 (&zhi)[8], const float2 (&recv)[8], float2 base,
                             float* mrow, int lane) {
  unfold_all512(zlo, recv, base, mrow, lane, std::make_integer_sequence<int, 8>{});
  if (lane == 0) {  // k = 256 pairs with itself; W_1024^256 = -i
    float mk, mm;
    folded_magnitudes(zhi[0], zhi[0], make_float2(0.f, -1.f), mk, mm);
    mrow[256] = mk;
  }
}

// =============================================================================
// n_fft = 2048 on the 512-point core: X[k] = E[k] + W_2048^k O[k], with E and O the spectra of
// the even and odd samples (two real sequences of 1024, each folded onto one 512-point FFT).
// Halves the registers and the transpose tile of the 32-points-per-lane form below.
// =============================================================================
// cos / sin of 2*pi*j/64 for j = 0..8
DMEL_HD constexpr float cos64(int j) {
  switch (j) {
    case 0: return 1.0f;
    case 1: return 0.99518472667219693f;
    case 2: return 0.98078528040323043f;
    case 3: return 0.95694033573220882f;
    case 4: return 0.92387953251128674f;
    case 5: return 0.88192126434835505f;
    case 6: return 0.83146961230254524f;
    case 7: return 0.77301045336273699f;
    default: return 0.70710678118654752f;
  }
}
DMEL_HD constexpr float sin64(int j) {
  switch (j) {
    case 0: return 0.0f;
    case 1: return 0.09801714032956060f;
    case 2: return 0.19509032201612825f;
    case 3: return 0.29028467725446233f;
    case 4: return 0.38268343236508978f;
    case 5: return 0.47139673682599764f;
    case 6: return 0.55557023301960218f;
    case 7: return 0.63439328416364549f;
    default: return 0.70710678118654752f;
  }
}

// Complex unfold of one bin pair of a real 1024-sequence from Z = FFT_512 of its fold:
// A = Z[k], Bm = Z[512-k], w = W_1024^k.  Returns ek2 = 2 S[k] and em2 = 2 conj(S[512-k]).
DMEL_HD void unfold_complex(float2 A, float2 Bm, float2 w, float2& ek2, float2& em2) {
  const float2 P = f2_fma(Bm, make_float2(1.f, -1.f), A);
  const float2 Q = f2_fma(f2_swap(A), make_float2(1.f, -1.f), f2_swap(Bm));
  ek2 = cfma(P, Q, w);
  em2 = twice_minus(P, ek2);
}
// Spectrum halves of one lane: k = lane + 32 j -> S[k] (k2) and conj S[512-k] (m2), both times 2.
struct HalfSpectrum {
  float2 k2[8], m2[8];
  float2 mid;  // lane 0 only: 2 S[256]
};
template <int J>
DMEL_HD void unfold_complex_one(const float2 (&zlo)[8], const float2 (&recv)[8], float2 base1024, HalfSpectrum& s) {
  unfold_complex(zlo[J], recv[J], mul_w32<J>(base1024), s.k2[J], s.m2[J]);
}
template <int... J>
DMEL_HD void unfold_complex_all(const float2 (&zlo)[8], const float2 (&recv)[8], float2 base1024, HalfSpectrum& s,
                                std::integer_sequence<int, J...>) {
  (unfold_complex_one<J>(zlo, recv, base1024, s), ...);
}
// zlo/zhi/recv as in unfold_store512; base1024 = W_1024^lane.
DMEL_HD void unfold_half_spectrum(const float2 (&zlo)[8], const float2 (&zhi)[8], const float2 (&recv)[8],
                                  float2 base1024, HalfSpectrum& s) {
  unfold_complex_all(zlo, recv, base1024, s, std::make_integer_sequence<int, 8>{});
  float2 unused;
  unfold_complex(zhi[0], zhi[0], make_float2(0.f, -1.f), s.mid, unused);  // k = 256 (meaningful on lane 0)
}

DMEL_HD float mag_from_twice(float2 x2) {
  const float2 sq = f2_mul(x2, x2);
  return fast_sqrt(fmaf(0.25f, sq.x + sq.y, kMagEps));
}
// Four magnitudes of the 2048-frame from one bin pair of E and O (all inputs are twice the value):
//   X[k]       = E[k] + w O[k]                 X[1024-k] = conj(E[k] - w O[k])
//   X[512-k]   = conj(Em + i w Om)             X[512+k]  = Em - i w Om       (Em = conj E[512-k] etc.)
// with w = W_2048^k.
template <int J>
DMEL_HD void combine2048_one(const HalfSpectrum& e, const HalfSpectrum& o, float2 base2048, float* mrow, int lane) {
  const float2 w = cmul_conj_cs(base2048, cos64(J), sin64(J));  // W_2048^{lane + 32 J} = base * W_64^J
  const float2 iw = make_float2(-w.y, w.x);
  const float2 xk = cfma(e.k2[J], o.k2[J], w);    // E + w O
  const float2 xm = cfma(e.m2[J], o.m2[J], iw);   // Em + i w Om
  const int k = lane + 32 * J;
  mrow[k] = mag_from_twice(xk);
  mrow[1024 - k] = mag_from_twice(twice_minus(e.k2[J], xk));
  mrow[512 - k] = mag_from_twice(xm);
  mrow[512 + k] = mag_from_twice(twice_minus(e.m2[J], xm));
}
template <int... J>
DMEL_HD void combine2048_all(const HalfSpectrum& e, const HalfSpectrum& o, float2 base2048, float* mrow, int lane,
                             std::integer_sequence<int, J...>) {
  (combine2048_one<J>(e, o, base2048, mrow, lane), ...);
}
// Writes the 1025 magnitudes of the frame into mrow (this lane's share). base2048 = W_2048^lane.
DMEL_HD void combine2048_store(const HalfSpectrum& e, const HalfSpectrum& o, float2 base2048, float* mrow, int lane) {
  combine2048_all(e, o, base2048, mrow, lane, std::make_integer_sequence<int, 8>{});
  if (lane == 0) {  // k = 256: W_2048^256 = (1 - i)/sqrt2
    const float2 t = cmul_conj_cs(o.mid, 0.70710678118654752f, 0.70710678118654752f);
    mrow[256] = mag_from_twice(cadd(e.mid, t));
    mrow[768] = mag_from_twice(csub(e.mid, t));
  }
}

// =============================================================================
// n_fft = 2048 : 1024-point complex FFT, 32 points per lane
// =============================================================================
// Pass 1, lane n2:  v[n1] = z[32*n1 + n2].  Leaves Y[n2][k1] * W_1024^{n2*k1}
// at tile[k1][n2].  tw[k1] = W_1024^{n2*k1}.
DMEL_HD void fft1024_pass1(float2 (&v)[32], const float2 (&tw)[32], float2* tile, int lane) {
  radix32(v);
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) {
    const float2 y = v[brev5(k1)];
    tile[k1 * kTilePitch + lane] = (k1 == 0) ? y : cmul(y, tw[k1]);
  }
}
// Pass 2, lane k1: leaves Z[k1 + 32*k2] in v[brev5(k2)].
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
// Z[1024-k] for k = lane + 32*k2 lives in lane (32-lane)&31, register
// brev5(mirror_slot1024(k2, that lane == 0)).
DMEL_HD constexpr int mirror_slot1024(int k2, bool sender_is_lane0) {
  return sender_is_lane0 ? ((32 - k2) & 31) : (31 - k2);
}

}  // namespace dmel
