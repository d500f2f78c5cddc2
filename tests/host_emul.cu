// CPU emulation of the warp FFTs in dmel_codec_b200/csrc/fft_core.cuh.
// Runs the SAME __host__ __device__ code lane by lane (shared-memory tile as a
// plain array, shuffles as array lookups) and checks spectra and magnitudes
// against a float64 DFT.  Built and run by tests/test_host_emul.py; no GPU.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../dmel_codec_b200/csrc/fft_core.cuh"

using namespace dmel;

static const double kTwoPi = 6.283185307179586476925286766559;

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

static double max_rel(const std::vector<double>& ref, const std::vector<float>& got, double floor_) {
  double worst = 0;
  for (size_t i = 0; i < ref.size(); ++i) {
    const double e = fabs(ref[i] - got[i]) / fmax(fabs(ref[i]), floor_);
    if (e > worst) worst = e;
  }
  return worst;
}

static float2 twiddle(int num, int den) {
  const double a = -kTwoPi * double(num % den) / den;
  return make_float2((float)cos(a), (float)sin(a));
}

// ---- n_fft = 1024: 512-point core -------------------------------------------
struct Warp512 {
  float2 zlo[32][8], zhi[32][8];
};

static void warp_fft512(const std::vector<float2>& z, Warp512& w, bool rebuilt_twiddles = false) {
  std::vector<float2> tile(kTile512);
  for (int lane = 0; lane < 32; ++lane) {
    float2 v[16], tw[16];
    for (int n1 = 0; n1 < 16; ++n1) v[n1] = z[32 * n1 + lane];
    for (int k1 = 0; k1 < 16; ++k1) tw[k1] = twiddle(lane * k1, 512);
    if (rebuilt_twiddles) fft512_pass1_pow(v, tw[1], tw[2], tw[4], tw[8], tile.data(), lane);
    else fft512_pass1(v, tw, tile.data(), lane);
  }
  static float2 g[32][16], send[32][8];
  for (int lane = 0; lane < 32; ++lane) {
    float2 v[16];
    fft512_pass2(v, tile.data(), lane);
    for (int i = 0; i < 16; ++i) g[lane][i] = v[i];
    float2 s[8];
    combine_send(v, lane >> 4, s);
    for (int j = 0; j < 8; ++j) send[lane][j] = s[j];
  }
  for (int lane = 0; lane < 32; ++lane) {
    float2 v[16], recv[8], zl[8], zh[8];
    for (int i = 0; i < 16; ++i) v[i] = g[lane][i];
    for (int j = 0; j < 8; ++j) recv[j] = send[lane ^ 16][j];
    combine_finish(v, recv, lane >> 4, zl, zh);
    for (int j = 0; j < 8; ++j) {
      w.zlo[lane][j] = zl[j];
      w.zhi[lane][j] = zh[j];
    }
  }
}

static void unfold512(const Warp512& w, std::vector<float>& mag) {
  mag.assign(513, -1.f);
  static float2 send[32][8];
  for (int lane = 0; lane < 32; ++lane) {
    float2 zl[8], zh[8], s[8];
    for (int j = 0; j < 8; ++j) { zl[j] = w.zlo[lane][j]; zh[j] = w.zhi[lane][j]; }
    mirror_send512(zl, zh, lane, s);
    for (int j = 0; j < 8; ++j) send[lane][j] = s[j];
  }
  for (int lane = 0; lane < 32; ++lane) {
    float2 zl[8], zh[8], recv[8];
    for (int j = 0; j < 8; ++j) { zl[j] = w.zlo[lane][j]; zh[j] = w.zhi[lane][j]; }
    const int src = mirror_lane512(lane);
    for (int j = 0; j < 8; ++j) recv[j] = send[src][j];
    unfold_store512(zl, zh, recv, twiddle(lane, 1024), mag.data(), lane);
  }
}

// ---- n_fft = 2048 on the 512-point core (even / odd split) -------------------
static void half_spectrum(const std::vector<float2>& z, HalfSpectrum out[32]) {
  static Warp512 w;
  warp_fft512(z, w, true);
  static float2 send[32][8];
  for (int lane = 0; lane < 32; ++lane) {
    float2 zl[8], zh[8], s[8];
    for (int j = 0; j < 8; ++j) { zl[j] = w.zlo[lane][j]; zh[j] = w.zhi[lane][j]; }
    mirror_send512(zl, zh, lane, s);
    for (int j = 0; j < 8; ++j) send[lane][j] = s[j];
  }
  for (int lane = 0; lane < 32; ++lane) {
    float2 zl[8], zh[8], recv[8];
    for (int j = 0; j < 8; ++j) { zl[j] = w.zlo[lane][j]; zh[j] = w.zhi[lane][j]; }
    for (int j = 0; j < 8; ++j) recv[j] = send[mirror_lane512(lane)][j];
    unfold_half_spectrum(zl, zh, recv, twiddle(lane, 1024), out[lane]);
  }
}

// ---- n_fft = 2048: 1024-point core ------------------------------------------
static void warp_fft1024(const std::vector<float2>& z, float2 regs[32][32]) {
  std::vector<float2> tile(kTile1024);
  for (int lane = 0; lane < 32; ++lane) {
    float2 v[32], tw[32];
    for (int n1 = 0; n1 < 32; ++n1) v[n1] = z[32 * n1 + lane];
    for (int k1 = 0; k1 < 32; ++k1) tw[k1] = twiddle(lane * k1, 1024);
    fft1024_pass1(v, tw, tile.data(), lane);
  }
  for (int lane = 0; lane < 32; ++lane) {
    float2 v[32];
    fft1024_pass2(v, tile.data(), lane);
    for (int s = 0; s < 32; ++s) regs[lane][s] = v[s];
  }
}

int main() {
  srand(7);
  auto rnd = []() { return (rand() / (double)RAND_MAX) * 2.0 - 1.0; };
  int bad = 0;

  // complex 512 core: Z[lane + 32 j] == zlo[lane][j], Z[lane + 32 j + 256] == zhi[lane][j]
  {
    std::vector<float2> z(512);
    for (auto& c : z) c = make_float2((float)rnd(), (float)rnd());
    static Warp512 w;
    warp_fft512(z, w);
    double worst = 0, norm = 0;
    for (int k = 0; k < 512; ++k) {
      double re = 0, im = 0;
      for (int i = 0; i < 512; ++i) {
        const double a = -kTwoPi * double((long long)k * i % 512) / 512.0;
        re += z[i].x * cos(a) - z[i].y * sin(a);
        im += z[i].x * sin(a) + z[i].y * cos(a);
      }
      const int kk = k % 256, lane = kk % 32, j = kk / 32;
      const float2 g = k < 256 ? w.zlo[lane][j] : w.zhi[lane][j];
      worst = fmax(worst, hypot(re - g.x, im - g.y));
      norm = fmax(norm, hypot(re, im));
    }
    printf("complex512  max_abs_err %.3e (max |Z| %.3f)\n", worst, norm);
    if (worst > 2e-5 * norm) { printf("FAIL complex 512 core\n"); ++bad; }
  }

  // real 1024-sample frames through fold + unfold, including a very quiet one; both twiddle variants
  for (int variant = 0; variant < 2; ++variant)
  for (double amp : {1.0, 1e-4}) {
    std::vector<double> x(1024);
    std::vector<float2> z(512);
    for (int i = 0; i < 1024; ++i) {
      const double win = 0.5 - 0.5 * cos(kTwoPi * i / 1024.0);
      x[i] = (double)(float)(amp * rnd() * win);
    }
    for (int i = 0; i < 512; ++i) z[i] = make_float2((float)x[2 * i], (float)x[2 * i + 1]);
    static Warp512 w;
    warp_fft512(z, w, variant == 1);
    std::vector<float> m;
    unfold512(w, m);
    const double e = max_rel(dft_mag(x), m, 1e-3 * amp);
    printf("real1024 %s amp %.0e max_rel_err %.3e\n", variant ? "rebuilt-tw" : "table-tw  ", amp, e);
    if (e > 2e-5) { printf("FAIL real 1024\n"); ++bad; }
  }

  // complex 1024 core + real 2048 unfold
  {
    std::vector<double> x(2048);
    std::vector<float2> z(1024);
    for (int i = 0; i < 2048; ++i) {
      const double win = 0.5 - 0.5 * cos(kTwoPi * i / 2048.0);
      x[i] = (double)(float)(rnd() * win);
    }
    for (int i = 0; i < 1024; ++i) z[i] = make_float2((float)x[2 * i], (float)x[2 * i + 1]);
    static float2 regs[32][32];
    warp_fft1024(z, regs);
    std::vector<float> m(1025);
    for (int lane = 0; lane < 32; ++lane) {
      for (int k2 = 0; k2 < 16; ++k2) {
        const int k = 32 * k2 + lane, sender = (32 - lane) & 31;
        const float2 bm = regs[sender][brev5(mirror_slot1024(k2, sender == 0))];
        folded_magnitudes(regs[lane][brev5(k2)], bm, twiddle(k, 2048), m[k], m[1024 - k]);
      }
      if (lane == 0) {
        float dummy;
        folded_magnitudes(regs[0][brev5(16)], regs[0][brev5(16)], twiddle(512, 2048), m[512], dummy);
      }
    }
    const double e = max_rel(dft_mag(x), m, 1e-3);
    printf("real2048 max_rel_err %.3e\n", e);
    if (e > 2e-5) { printf("FAIL real 2048\n"); ++bad; }
  }
  // real 2048 through two 512-point FFTs (even / odd samples)
  for (double amp : {1.0, 1e-4}) {
    std::vector<double> x(2048);
    for (int i = 0; i < 2048; ++i) {
      const double win = 0.5 - 0.5 * cos(kTwoPi * i / 2048.0);
      x[i] = (double)(float)(amp * rnd() * win);
    }
    std::vector<float2> ze(512), zo(512);
    for (int n = 0; n < 512; ++n) {
      ze[n] = make_float2((float)x[4 * n], (float)x[4 * n + 2]);
      zo[n] = make_float2((float)x[4 * n + 1], (float)x[4 * n + 3]);
    }
    static HalfSpectrum e[32], o[32];
    half_spectrum(ze, e);
    half_spectrum(zo, o);
    std::vector<float> m(1025, -1.f);
    for (int lane = 0; lane < 32; ++lane) combine2048_store(e[lane], o[lane], twiddle(lane, 2048), m.data(), lane);
    const double err = max_rel(dft_mag(x), m, 1e-3 * amp);
    printf("real2048 even/odd amp %.0e max_rel_err %.3e\n", amp, err);
    if (err > 2e-5) { printf("FAIL real 2048 even/odd\n"); ++bad; }
  }
  printf(bad ? "HOST_EMUL FAIL\n" : "HOST_EMUL OK\n");
  return bad;
}
