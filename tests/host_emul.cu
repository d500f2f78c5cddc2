// CPU emulation of the warp FFT in dmel_codec_b200/csrc/fft_core.cuh.
// Runs the SAME __host__ __device__ code lane by lane (shared-memory tile as a
// plain array, shuffles as array lookups) and checks magnitudes against a
// float64 DFT.  Built and run by tests/test_host_emul.py; needs no GPU.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../dmel_codec_b200/csrc/fft_core.cuh"

using namespace dmel;

static const double kTwoPi = 6.283185307179586476925286766559;

static void stage_twiddles(std::vector<float2>& tw) {
  tw.resize(32 * 32);
  for (int k1 = 0; k1 < 32; ++k1)
    for (int n2 = 0; n2 < 32; ++n2) {
      const double a = -kTwoPi * double((k1 * n2) % 1024) / 1024.0;
      tw[k1 * 32 + n2] = make_float2((float)cos(a), (float)sin(a));
    }
}

// Z = FFT_1024(z) through the two register passes; regs[lane][slot] as left by pass 2
static void warp_fft(const std::vector<float2>& z, float2 regs[32][32]) {
  std::vector<float2> stw;
  stage_twiddles(stw);
  std::vector<float2> tile(kTileFloat2);
  for (int lane = 0; lane < 32; ++lane) {
    float2 v[32], tw[32];
    for (int n1 = 0; n1 < 32; ++n1) v[n1] = z[32 * n1 + lane];
    for (int k1 = 0; k1 < 32; ++k1) tw[k1] = stw[k1 * 32 + lane];
    fft1024_pass1(v, tw, tile.data(), lane);
  }
  for (int lane = 0; lane < 32; ++lane) {
    float2 v[32];
    fft1024_pass2(v, tile.data(), lane);
    for (int s = 0; s < 32; ++s) regs[lane][s] = v[s];
  }
}

static float2 mirror_of(float2 regs[32][32], int lane, int k2) {
  const int sender = (32 - lane) & 31;
  return regs[sender][brev5(mirror_slot(k2, sender == 0))];
}

static double max_rel(const std::vector<double>& ref, const std::vector<float>& got, double floor_) {
  double worst = 0;
  for (size_t i = 0; i < ref.size(); ++i) {
    const double e = fabs(ref[i] - got[i]) / fmax(fabs(ref[i]), floor_);
    if (e > worst) worst = e;
  }
  return worst;
}

static std::vector<double> dft_mag(const std::vector<double>& x) {
  const int n = (int)x.size();
  std::vector<double> mag(n / 2 + 1);
  for (int k = 0; k <= n / 2; ++k) {
    double re = 0, im = 0;
    for (int i = 0; i < n; ++i) {
      const double a = -kTwoPi * double((long long)k * i % n) / n;
      re += x[i] * cos(a);
      im += x[i] * sin(a);
    }
    mag[k] = sqrt(re * re + im * im + 1e-9);
  }
  return mag;
}

int main() {
  srand(7);
  auto rnd = []() { return (rand() / (double)RAND_MAX) * 2.0 - 1.0; };
  int bad = 0;

  // ---- complex core: Z[lane + 32*k2] == regs[lane][brev5(k2)] --------------
  {
    std::vector<float2> z(1024);
    std::vector<double> zr(1024), zi(1024);
    for (int i = 0; i < 1024; ++i) {
      zr[i] = rnd();
      zi[i] = rnd();
      z[i] = make_float2((float)zr[i], (float)zi[i]);
      zr[i] = z[i].x;
      zi[i] = z[i].y;
    }
    static float2 regs[32][32];
    warp_fft(z, regs);
    double worst = 0, norm = 0;
    for (int k = 0; k < 1024; ++k) {
      double re = 0, im = 0;
      for (int i = 0; i < 1024; ++i) {
        const double a = -kTwoPi * double((long long)k * i % 1024) / 1024.0;
        re += zr[i] * cos(a) - zi[i] * sin(a);
        im += zr[i] * sin(a) + zi[i] * cos(a);
      }
      const float2 g = regs[k % 32][brev5(k / 32)];
      worst = fmax(worst, hypot(re - g.x, im - g.y));
      norm = fmax(norm, hypot(re, im));
    }
    printf("complex1024 max_abs_err %.3e (max |Z| %.3f)\n", worst, norm);
    if (worst > 2e-4 * norm) { printf("FAIL complex core\n"); ++bad; }
  }

  // ---- packed pair (n_fft = 1024): two real frames per FFT -----------------
  {
    std::vector<double> xa(1024), xb(1024);
    std::vector<float2> z(1024);
    for (int i = 0; i < 1024; ++i) {
      const double w = 0.5 - 0.5 * cos(kTwoPi * i / 1024.0);
      z[i] = make_float2((float)(rnd() * w), (float)(0.01 * rnd() * w));
      xa[i] = z[i].x;
      xb[i] = z[i].y;
    }
    static float2 regs[32][32];
    warp_fft(z, regs);
    std::vector<float> ma(513), mb(513);
    for (int lane = 0; lane < 32; ++lane) {
      for (int k2 = 0; k2 < 16; ++k2)
        packed_pair_magnitudes(regs[lane][brev5(k2)], mirror_of(regs, lane, k2), ma[32 * k2 + lane], mb[32 * k2 + lane]);
      if (lane == 0) packed_pair_magnitudes(regs[0][brev5(16)], regs[0][brev5(16)], ma[512], mb[512]);
    }
    const double ea = max_rel(dft_mag(xa), ma, 1e-3), eb = max_rel(dft_mag(xb), mb, 1e-3);
    printf("packed1024 max_rel_err frame_a %.3e frame_b %.3e\n", ea, eb);
    if (ea > 1e-4 || eb > 1e-2) { printf("FAIL packed pair\n"); ++bad; }  // frame_b is 100x quieter than its pair
  }

  // ---- folded (n_fft = 2048): one real frame per FFT ------------------------
  {
    std::vector<double> x(2048);
    std::vector<float2> z(1024), fold(513);
    for (int i = 0; i < 2048; ++i) {
      const double w = 0.5 - 0.5 * cos(kTwoPi * i / 2048.0);
      x[i] = (double)(float)(rnd() * w);
    }
    for (int i = 0; i < 1024; ++i) z[i] = make_float2((float)x[2 * i], (float)x[2 * i + 1]);
    for (int k = 0; k <= 512; ++k) {
      const double a = -kTwoPi * k / 2048.0;
      fold[k] = make_float2((float)cos(a), (float)sin(a));
    }
    static float2 regs[32][32];
    warp_fft(z, regs);
    std::vector<float> m(1025);
    for (int lane = 0; lane < 32; ++lane) {
      for (int k2 = 0; k2 < 16; ++k2) {
        const int k = 32 * k2 + lane;
        folded_magnitudes(regs[lane][brev5(k2)], mirror_of(regs, lane, k2), fold[k], m[k], m[1024 - k]);
      }
      if (lane == 0) {
        float dummy;
        folded_magnitudes(regs[0][brev5(16)], regs[0][brev5(16)], fold[512], m[512], dummy);
      }
    }
    const double e = max_rel(dft_mag(x), m, 1e-3);
    printf("folded2048 max_rel_err %.3e\n", e);
    if (e > 1e-4) { printf("FAIL folded\n"); ++bad; }
  }
  printf(bad ? "HOST_EMUL FAIL\n" : "HOST_EMUL OK\n");
  return bad;
}
