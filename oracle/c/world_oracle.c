/*
 * CPU oracle in C (TEST INFRASTRUCTURE and CPU baseline, not product code).
 *
 * Plain-C, double-precision, single-threaded-per-utterance restatement of the WORLD / SPTK routines IdiapTTS reaches
 * through pyworld / pysptk on the analysis half of the WORLD feature path (the same algorithms as oracle/world_np.py and
 * oracle/sptk_np.py, which are pinned against the reference's golden fixtures; tests/test_oracle_c.py pins this file
 * against both).  Neither pyworld (mmorise/World) nor pysptk (r9y9/SPTK) sources exist under /root/reference, so nothing
 * here is compiled from reference sources: the published algorithms are restated.
 *
 *   cheaptrick        WORLD cheaptrick.cpp            (reference call site WorldFeatLabelGen.py:792)
 *   d4c_coarse        WORLD d4c.cpp incl. LoveTrain   (:792), bap = WORLD codec.cpp CodeAperiodicity (:805)
 *   mcep              SPTK mcep.c / freqt / frqtr / (Toeplitz+Hankel solve)   (AudioProcessing.py:146)
 *   lf0 / vuv         WorldFeatLabelGen.py:798-802 + misc/utils.py:40-86
 *   oracle_extract    the per-utterance body of WorldFeatLabelGen.gen_data (:996-1013) with a cached F0 track
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define K_PI 3.1415926535897932384
#define K_SAFE 1e-12
#define K_EPS 2.2204460492503131e-16
#define K_LOG2 0.69314718055994529
#define K_DEFAULT_F0 500.0
#define K_FREQ_INTERVAL 3000.0
#define K_UPPER 15000.0
#define K_FLOOR_D4C 47.0

static int mround(double x) { return x > 0 ? (int)(x + 0.5) : (int)(x - 0.5); }
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

int oracle_cheaptrick_fft_size(int fs, double f0_floor) { return (int)pow(2.0, 1.0 + (int)(log(3.0 * fs / f0_floor + 1) / K_LOG2)); }
int oracle_num_aperiodicities(int fs) {
  double v = fs / 2.0 - K_FREQ_INTERVAL;
  if (v > K_UPPER) v = K_UPPER;
  return (int)(v / K_FREQ_INTERVAL);
}
int oracle_d4c_fft_size(int fs) { return (int)pow(2.0, 1.0 + (int)(log(4.0 * fs / K_FLOOR_D4C + 1) / K_LOG2)); }

/* ---- FFT: iterative radix-2 complex FFT with a per-size twiddle cache; real transforms through the half-size trick ---- */
typedef struct {
  int n;
  double* wr;
  double* wi;
  int* rev;
} fft_plan;

static fft_plan* plan_create(int n) {
  fft_plan* p = (fft_plan*)malloc(sizeof(fft_plan));
  p->n = n;
  p->wr = (double*)malloc(sizeof(double) * n);
  p->wi = (double*)malloc(sizeof(double) * n);
  p->rev = (int*)malloc(sizeof(int) * n);
  int bits = 0;
  while ((1 << bits) < n) ++bits;
  for (int i = 0; i < n; ++i) {
    p->wr[i] = cos(-2.0 * K_PI * i / n);
    p->wi[i] = sin(-2.0 * K_PI * i / n);
    int r = 0;
    for (int b = 0; b < bits; ++b)
      if (i & (1 << b)) r |= 1 << (bits - 1 - b);
    p->rev[i] = r;
  }
  return p;
}
static void plan_free(fft_plan* p) {
  free(p->wr); free(p->wi); free(p->rev); free(p);
}
/* in-place forward complex FFT of length p->n */
static void cfft(const fft_plan* p, double* re, double* im) {
  const int n = p->n;
  for (int i = 0; i < n; ++i) {
    const int j = p->rev[i];
    if (j > i) {
      double t = re[i]; re[i] = re[j]; re[j] = t;
      t = im[i]; im[i] = im[j]; im[j] = t;
    }
  }
  for (int len = 2; len <= n; len <<= 1) {
    const int half = len >> 1, step = n / len;
    for (int s = 0; s < n; s += len) {
      for (int k = 0; k < half; ++k) {
        const double wr = p->wr[k * step], wi = p->wi[k * step];
        const int a = s + k, b = a + half;
        const double xr = re[b] * wr - im[b] * wi, xi = re[b] * wi + im[b] * wr;
        re[b] = re[a] - xr; im[b] = im[a] - xi;
        re[a] += xr; im[a] += xi;
      }
    }
  }
}
typedef struct {
  int n;           /* real length */
  fft_plan* half;  /* complex plan of n/2 */
  fft_plan* full;  /* twiddles exp(-2 pi i k / n) */
  double *zr, *zi;
} rfft_plan;
static rfft_plan* rplan_create(int n) {
  rfft_plan* p = (rfft_plan*)malloc(sizeof(rfft_plan));
  p->n = n;
  p->half = plan_create(n / 2);
  p->full = plan_create(n);
  p->zr = (double*)malloc(sizeof(double) * n / 2);
  p->zi = (double*)malloc(sizeof(double) * n / 2);
  return p;
}
static void rplan_free(rfft_plan* p) {
  plan_free(p->half); plan_free(p->full); free(p->zr); free(p->zi); free(p);
}
/* X[0..n/2] of the real sequence x[0..n) */
static void rfft(rfft_plan* p, const double* x, double* xr, double* xi) {
  const int n = p->n, m = n / 2;
  for (int i = 0; i < m; ++i) { p->zr[i] = x[2 * i]; p->zi[i] = x[2 * i + 1]; }
  cfft(p->half, p->zr, p->zi);
  for (int k = 0; k <= m; ++k) {
    const int a = k & (m - 1), b = (m - k) & (m - 1);
    const double ar = p->zr[a], ai = p->zi[a], br = p->zr[b], bi = -p->zi[b];
    const double er = 0.5 * (ar + br), ei = 0.5 * (ai + bi);
    const double dr = 0.5 * (ar - br), di = 0.5 * (ai - bi);
    const double or_ = di, oi = -dr; /* (a - b) / (2i) */
    const double wr = (k == m) ? -1.0 : p->full->wr[k], wi = (k == m) ? 0.0 : p->full->wi[k];
    xr[k] = er + wr * or_ - wi * oi;
    xi[k] = ei + wr * oi + wi * or_;
  }
}

/* ---- WORLD common helpers -------------------------------------------------------------------------------------------- */
static double interp1q_at(double x0, double dx, const double* y, int ylen, double xi) {
  const double pos = (xi - x0) / dx;
  const int base = (int)pos;
  const double frac = pos - base;
  const double dy = (base + 1 < ylen) ? y[base + 1] - y[base] : 0.0;
  return y[base] + dy * frac;
}
static void dc_correction(double* p, double f0, int fs, int n, double* tmp) {
  const int upper = 2 + (int)(f0 * n / fs);
  for (int i = 0; i < upper - 1; ++i) tmp[i] = interp1q_at(f0, -(double)fs / n, p, upper + 1, (double)i * fs / n);
  for (int i = 0; i < upper - 1; ++i) p[i] += tmp[i];
}
/* out may alias in; seg: scratch of n/2 + 2 b + 1 */
static void linear_smoothing(const double* in, double width, int fs, int n, double* out, double* seg) {
  const int b = (int)(width * n / fs) + 1, h = n / 2, L = h + 2 * b + 1;
  for (int i = 0; i < L; ++i) {
    double v;
    if (i < b) v = in[b - i];
    else if (i < h + b) v = in[i - b];
    else v = in[h - (i - (h + b))];
    seg[i] = v * fs / n + (i ? seg[i - 1] : 0.0);
  }
  const double origin = -(b - 0.5) * fs / n, dfi = (double)fs / n;
  for (int k = 0; k <= h; ++k) {
    const double fax = (double)k / n * fs - width / 2.0;
    const double lo = interp1q_at(origin, dfi, seg, L, fax);
    const double hi = interp1q_at(origin, dfi, seg, L, fax + width);
    out[k] = (hi - lo) / width;
  }
}
static double sample_at(const double* x, int xlen, int idx) { return x[imax(0, imin(xlen - 1, idx))]; }

/* ---- CheapTrick ------------------------------------------------------------------------------------------------------------ */
int oracle_cheaptrick(const double* x, int xlen, int fs, const double* f0, const double* t, int T, int fft_size, double q1,
                      double* sp) {
  const int n = fft_size, h = n / 2, K = h + 1;
  rfft_plan* rp = rplan_create(n);
  double* buf = (double*)calloc(n, sizeof(double));
  double* win = (double*)malloc(sizeof(double) * n);
  double* xr = (double*)malloc(sizeof(double) * K);
  double* xi = (double*)malloc(sizeof(double) * K);
  double* p = (double*)malloc(sizeof(double) * K);
  double* tmp = (double*)malloc(sizeof(double) * (K + 2 * K + 4));
  const double floor_f0 = 3.0 * fs / (n - 3.0);
  for (int fr = 0; fr < T; ++fr) {
    const double cur = f0[fr] <= floor_f0 ? K_DEFAULT_F0 : f0[fr];
    const int half = mround(1.5 * fs / cur), wl = 2 * half + 1;
    const int origin = mround(t[fr] * fs + 0.001);
    double avg = 0.0;
    for (int i = 0; i < wl; ++i) {
      const double position = (i - half) / 1.5 / fs;
      win[i] = 0.5 * cos(K_PI * position * cur) + 0.5;
      avg += win[i] * win[i];
    }
    avg = sqrt(avg);
    double sw = 0.0, swin = 0.0;
    for (int i = 0; i < wl; ++i) {
      win[i] /= avg;
      buf[i] = sample_at(x, xlen, origin + i - half) * win[i];
      sw += buf[i];
      swin += win[i];
    }
    const double coef = sw / swin;
    for (int i = 0; i < wl; ++i) buf[i] -= win[i] * coef;
    for (int i = wl; i < n; ++i) buf[i] = 0.0;
    rfft(rp, buf, xr, xi);
    for (int k = 0; k < K; ++k) p[k] = xr[k] * xr[k] + xi[k] * xi[k];
    dc_correction(p, cur, fs, n, tmp);
    linear_smoothing(p, cur * 2.0 / 3.0, fs, n, p, tmp);
    for (int k = 0; k < K; ++k) buf[k] = log(p[k] + K_EPS);
    for (int k = 1; k < h; ++k) buf[n - k] = buf[k];
    rfft(rp, buf, xr, xi);
    for (int k = 0; k < K; ++k) {
      const double quef = (double)k / fs;
      const double smooth = k == 0 ? 1.0 : sin(K_PI * cur * quef) / (K_PI * cur * quef);
      const double comp = (1.0 - 2.0 * q1) + 2.0 * q1 * cos(2.0 * K_PI * quef * cur);
      buf[k] = xr[k] * smooth * comp / n;
    }
    for (int k = 1; k < h; ++k) buf[n - k] = buf[k];
    rfft(rp, buf, xr, xi);
    for (int k = 0; k < K; ++k) sp[(size_t)fr * K + k] = exp(xr[k]);
  }
  free(buf); free(win); free(xr); free(xi); free(p); free(tmp);
  rplan_free(rp);
  return 0;
}

/* ---- D4C ------------------------------------------------------------------------------------------------------------------- */
static void d4c_window(const double* x, int xlen, int fs, double f0, double pos, int blackman, double ratio, double* out, int n) {
  const int half = mround(ratio * fs / f0 / 2.0), wl = imin(n, 2 * half + 1);
  const int origin = mround(pos * fs + 0.001);
  double sw = 0.0, swin = 0.0;
  double* win = out + n; /* caller provides 2 n doubles */
  for (int i = 0; i < wl; ++i) {
    const double position = (2.0 * (i - half) / ratio) / fs;
    const double c = cos(K_PI * position * f0);
    win[i] = blackman ? 0.42 + 0.5 * c + 0.08 * cos(K_PI * position * f0 * 2) : 0.5 * c + 0.5;
    out[i] = sample_at(x, xlen, origin + i - half) * win[i];
    sw += out[i];
    swin += win[i];
  }
  const double coef = sw / swin;
  for (int i = 0; i < wl; ++i) out[i] -= win[i] * coef;
  for (int i = wl; i < n; ++i) out[i] = 0.0;
}
static int cmp_double(const void* a, const void* b) {
  const double x = *(const double*)a, y = *(const double*)b;
  return (x > y) - (x < y);
}
int oracle_d4c_coarse(const double* x, int xlen, int fs, const double* f0, const double* t, int T, double threshold, double* coarse,
                      unsigned char* voiced) {
  const int n = oracle_d4c_fft_size(fs), h = n / 2, K = h + 1;
  const int nap = oracle_num_aperiodicities(fs);
  const int wl = (int)(K_FREQ_INTERVAL * n / fs) * 2 + 1, hw = wl / 2;
  const int boundary = mround(n * 8.0 / wl);
  rfft_plan* rp = rplan_create(n);
  double* nutt = (double*)malloc(sizeof(double) * wl);
  for (int i = 0; i < wl; ++i) {
    const double tmp = i / (wl - 1.0);
    nutt[i] = 0.355768 - 0.487396 * cos(2.0 * K_PI * tmp) + 0.144232 * cos(4.0 * K_PI * tmp) - 0.012604 * cos(6.0 * K_PI * tmp);
  }
  double* buf = (double*)malloc(sizeof(double) * 2 * n);
  double* w2 = (double*)malloc(sizeof(double) * n);
  double *ar = (double*)malloc(sizeof(double) * K), *ai = (double*)malloc(sizeof(double) * K);
  double *br = (double*)malloc(sizeof(double) * K), *bi = (double*)malloc(sizeof(double) * K);
  double* sc = (double*)malloc(sizeof(double) * K);
  double* sps = (double*)malloc(sizeof(double) * K);
  double* gd = (double*)malloc(sizeof(double) * K);
  double* sm = (double*)malloc(sizeof(double) * K);
  double* tmp = (double*)malloc(sizeof(double) * (3 * K + 8));
  const int b0 = (int)ceil(100.0 * n / fs), b1 = imin(h, (int)ceil(4000.0 * n / fs)), b2 = imin(h, (int)ceil(7900.0 * n / fs));
  for (int fr = 0; fr < T; ++fr) {
    voiced[fr] = 0;
    for (int b = 0; b < nap; ++b) coarse[(size_t)fr * nap + b] = 0.0;
    if (f0[fr] == 0.0) continue;
    /* LoveTrain */
    d4c_window(x, xlen, fs, fmax(f0[fr], 40.0), t[fr], 1, 3.0, buf, n);
    rfft(rp, buf, ar, ai);
    double c1 = 0.0, c2 = 0.0;
    for (int k = b0 + 1; k <= b2; ++k) {
      c2 += ar[k] * ar[k] + ai[k] * ai[k];
      if (k == b1) c1 = c2;
    }
    if (!(c1 / c2 > threshold)) continue;
    voiced[fr] = 1;
    const double f = fmax(K_FLOOR_D4C, f0[fr]);
    /* static centroid */
    for (int side = 0; side < 2; ++side) {
      const double tc = side == 0 ? t[fr] - 0.25 / f : t[fr] + 0.25 / f;
      d4c_window(x, xlen, fs, f, tc, 1, 4.0, buf, n);
      const int nl = imin(n, mround(2.0 * fs / f) * 2 + 1);
      double pw = 0.0;
      for (int i = 0; i < nl; ++i) pw += buf[i] * buf[i];
      pw = sqrt(pw);
      for (int i = 0; i < nl; ++i) buf[i] /= pw;
      for (int i = 0; i < n; ++i) w2[i] = buf[i] * (i + 1.0);
      rfft(rp, buf, ar, ai);
      rfft(rp, w2, br, bi);
      for (int k = 0; k < K; ++k) {
        const double cen = br[k] * ar[k] + ai[k] * bi[k];
        sc[k] = side == 0 ? cen : sc[k] + cen;
      }
    }
    dc_correction(sc, f, fs, n, tmp);
    /* smoothed power spectrum */
    d4c_window(x, xlen, fs, f, t[fr], 0, 4.0, buf, n);
    rfft(rp, buf, ar, ai);
    for (int k = 0; k < K; ++k) sps[k] = ar[k] * ar[k] + ai[k] * ai[k];
    dc_correction(sps, f, fs, n, tmp);
    linear_smoothing(sps, f, fs, n, sps, tmp);
    /* static group delay */
    for (int k = 0; k < K; ++k) gd[k] = sc[k] / sps[k];
    linear_smoothing(gd, f / 2.0, fs, n, gd, tmp);
    linear_smoothing(gd, f, fs, n, sm, tmp);
    for (int k = 0; k < K; ++k) gd[k] -= sm[k];
    /* coarse aperiodicity */
    for (int b = 0; b < nap; ++b) {
      const int center = (int)(K_FREQ_INTERVAL * (b + 1) * n / fs);
      for (int i = 0; i < n; ++i) buf[i] = i < wl ? gd[center - hw + i] * nutt[i] : 0.0;
      rfft(rp, buf, ar, ai);
      for (int k = 0; k < K; ++k) sm[k] = ar[k] * ar[k] + ai[k] * ai[k];
      qsort(sm, K, sizeof(double), cmp_double);
      for (int k = 1; k < K; ++k) sm[k] += sm[k - 1];
      double v = 10.0 * log10(sm[h - boundary - 1] / sm[h]) + (f - 100.0) / 50.0;
      coarse[(size_t)fr * nap + b] = v < 0.0 ? v : 0.0;
    }
  }
  free(nutt); free(buf); free(w2); free(ar); free(ai); free(br); free(bi); free(sc); free(sps); free(gd); free(sm); free(tmp);
  rplan_free(rp);
  return 0;
}

/* interp1 of the coarse aperiodicity (dB) at frequency f: knots 0, 3000, .., 3000 nap, fs/2; values -60, coarse.., -1e-12 */
static double coarse_db_at(const double* coarse, int nap, double fs_half, double f) {
  int k = (int)(f / K_FREQ_INTERVAL) + 1;
  if (k > nap + 1) k = nap + 1;
  const double xl = (k - 1) * K_FREQ_INTERVAL, xr = (k == nap + 1) ? fs_half : k * K_FREQ_INTERVAL;
  const double yl = (k - 1 == 0) ? -60.0 : coarse[k - 2], yr = (k == nap + 1) ? -K_SAFE : coarse[k - 1];
  return yl + (f - xl) / (xr - xl) * (yr - yl);
}
/* pyworld.code_aperiodicity(d4c(...)) from the coarse values */
void oracle_bap_from_coarse(const double* coarse, const unsigned char* voiced, int T, int fs, int fft_size, double* bap) {
  const int nap = oracle_num_aperiodicities(fs), K = fft_size / 2 + 1;
  for (int fr = 0; fr < T; ++fr) {
    for (int b = 0; b < nap; ++b) {
      const double pos = (K_FREQ_INTERVAL * (b + 1.0)) / ((double)fs / fft_size);
      const int base = (int)pos;
      const double frac = pos - base;
      double y[2];
      for (int q = 0; q < 2; ++q) {
        double a = 1.0 - K_SAFE;
        if (voiced[fr]) a = pow(10.0, coarse_db_at(coarse + (size_t)fr * nap, nap, fs / 2.0, (double)(base + q) * fs / fft_size) / 20.0);
        y[q] = 20.0 * log10(a);
      }
      const double dy = (base + 1 < K) ? y[1] - y[0] : 0.0;
      bap[(size_t)fr * nap + b] = y[0] + dy * frac;
    }
  }
}

/* ---- SPTK mcep --------------------------------------------------------------------------------------------------------------- */
static void freqt(const double* c1, int m1, double* c2, int m2, double a, double* d) {
  const double b = 1.0 - a * a;
  double* g = c2;
  for (int j = 0; j <= m2; ++j) g[j] = 0.0;
  for (int i = -m1; i <= 0; ++i) {
    if (0 <= m2) { d[0] = g[0]; g[0] = c1[-i] + a * d[0]; }
    if (1 <= m2) { d[1] = g[1]; g[1] = b * d[0] + a * d[1]; }
    for (int j = 2; j <= m2; ++j) { d[j] = g[j]; g[j] = d[j - 1] + a * (d[j] - g[j - 1]); }
  }
}
static void frqtr(const double* c1, int m1, double* c2, int m2, double a, double* d) {
  double* g = c2;
  for (int j = 0; j <= m2; ++j) g[j] = 0.0;
  for (int i = -m1; i <= 0; ++i) {
    if (0 <= m2) { d[0] = g[0]; g[0] = c1[-i]; }
    for (int j = 1; j <= m2; ++j) { d[j] = g[j]; g[j] = d[j - 1] + a * (d[j] - g[j - 1]); }
  }
}
/* solves M x = b, M symmetric positive definite n x n (row-major, destroyed): Cholesky.  Returns -1 when not SPD. */
static int chol_solve(double* M, double* b, int n) {
  for (int j = 0; j < n; ++j) {
    double s = M[j * n + j];
    for (int k = 0; k < j; ++k) s -= M[j * n + k] * M[j * n + k];
    if (!(s > 0.0)) return -1;
    const double d = sqrt(s);
    M[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double v = M[i * n + j];
      for (int k = 0; k < j; ++k) v -= M[i * n + k] * M[j * n + k];
      M[i * n + j] = v / d;
    }
  }
  for (int i = 0; i < n; ++i) {
    double v = b[i];
    for (int k = 0; k < i; ++k) v -= M[i * n + k] * b[k];
    b[i] = v / M[i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double v = b[i];
    for (int k = i + 1; k < n; ++k) v -= M[k * n + i] * b[k];
    b[i] = v / M[i * n + i];
  }
  return 0;
}
/* amp [T x K] amplitude (is_power = 0) or power (1) spectrum -> mc [T x (m+1)]; returns 0, 3 (solve failed) or 4 (zero periodogram) */
int oracle_mcep(const double* amp, int is_power, int T, int fft_size, int m, double a, int itr1, int itr2, double dd, double eps,
                double* mc, int* iters) {
  const int n = fft_size, f2 = n / 2, K = f2 + 1, m2 = 2 * m;
  rfft_plan* rp = rplan_create(n);
  double* per = (double*)malloc(sizeof(double) * K);
  double* c = (double*)malloc(sizeof(double) * n);
  double* xr = (double*)malloc(sizeof(double) * K);
  double* xi = (double*)malloc(sizeof(double) * K);
  double* d = (double*)malloc(sizeof(double) * (n + 2));
  double* rt = (double*)malloc(sizeof(double) * (m2 + 2));
  double* M = (double*)malloc(sizeof(double) * (m + 1) * (m + 1));
  double* bvec = (double*)malloc(sizeof(double) * (m + 1));
  double* al = (double*)malloc(sizeof(double) * (m + 1));
  al[0] = 1.0;
  for (int i = 1; i <= m; ++i) al[i] = -a * al[i - 1];
  int rc = 0;
  for (int fr = 0; fr < T && rc == 0; ++fr) {
    const double* x = amp + (size_t)fr * K;
    double* out = mc + (size_t)fr * (m + 1);
    for (int k = 0; k < K; ++k) {
      per[k] = (is_power ? x[k] : x[k] * x[k]) + eps;
      if (!(per[k] > 0.0)) rc = 4;
    }
    if (rc) break;
    /* c = IFFT(log per): the spectrum is real and even, so the inverse transform is a forward real FFT scaled by 1/n */
    for (int k = 0; k < K; ++k) c[k] = log(per[k]);
    for (int k = 1; k < f2; ++k) c[n - k] = c[k];
    rfft(rp, c, xr, xi);
    for (int k = 0; k < K; ++k) c[k] = xr[k] / n;
    c[0] /= 2.0;
    c[f2] /= 2.0;
    freqt(c, f2, out, m, a, d);
    double s = c[0];
    int it = 0;
    for (int j = 1; j <= itr2; ++j) {
      it = j;
      freqt(out, m, c, f2, -a, d);
      for (int k = f2 + 1; k < n; ++k) c[k] = 0.0;
      rfft(rp, c, xr, xi); /* xr = Re FFT[c] */
      for (int k = 0; k < K; ++k) c[k] = per[k] / exp(xr[k] + xr[k]);
      for (int k = 1; k < f2; ++k) c[n - k] = c[k];
      rfft(rp, c, xr, xi);
      for (int k = 0; k < K; ++k) c[k] = xr[k] / n; /* r = IFFT of a real even sequence */
      frqtr(c, f2, rt, m2, a, d);
      const double tt = rt[0];
      if (j >= itr1) {
        if (fabs((tt - s) / tt) < dd) break;
        s = tt;
      }
      for (int i = 0; i <= m; ++i) {
        bvec[i] = rt[i] - al[i];
        for (int k = 0; k <= m; ++k) M[i * (m + 1) + k] = rt[abs(i - k)] + rt[i + k];
      }
      if (chol_solve(M, bvec, m + 1)) { rc = 3; break; }
      for (int i = 0; i <= m; ++i) out[i] += bvec[i];
    }
    if (iters) iters[fr] = it;
  }
  free(per); free(c); free(xr); free(xi); free(d); free(rt); free(M); free(bvec); free(al);
  rplan_free(rp);
  return rc;
}

/* ---- lf0 / vuv (WorldFeatLabelGen.py:798-802, misc/utils.py:40-86), float32 arithmetic like the reference -------------------- */
void oracle_lf0_vuv(const double* f0, int n, double thr, float lf0_zero, float* lf0, float* vuv) {
  const float log_thr = (float)log(thr);
  for (int i = 0; i < n; ++i) {
    const float x = (float)(f0[i] > 1e-10 ? f0[i] : 1e-10);
    float l = logf(x);
    if (l <= log_thr) l = lf0_zero;
    lf0[i] = l;
    vuv[i] = l > 0.f ? 1.f : 0.f;
  }
  float last = 0.f;
  int i = 0;
  while (i < n) {
    if (lf0[i] <= 0.f) {
      int j = i + 1;
      for (int jj = i + 1; jj < n; ++jj) {
        j = jj;
        if (lf0[jj] > 0.f) break;
      }
      if (j < n - 1) {
        if (last > 0.f) {
          const float prev = lf0[i - 1];
          const float step = (lf0[j] - prev) / (float)(j - i);
          for (int k = i; k < j; ++k) {
            volatile float pr = step * (float)(k - i + 1); /* no fused multiply-add */
            lf0[k] = prev + pr;
          }
        } else {
          for (int k = i; k < j; ++k) lf0[k] = lf0[j];
        }
        last = lf0[j - 1];
        i = j;
      } else {
        for (int k = i; k < n; ++k) lf0[k] = last;
        break;
      }
    } else {
      last = lf0[i];
      ++i;
    }
  }
}

/* ---- one utterance of WorldFeatLabelGen.gen_data with a cached F0 track ------------------------------------------------------ */
/* wave: int16 PCM; feats [T x (D + 2 + nap)] float32 = [mcep | lf0 | vuv | bap] */
int oracle_extract(const int16_t* wave, int xlen, int fs, const double* f0, int T, double preemphasis, int num_coded_sps, double alpha,
                   float* feats) {
  const int n = oracle_cheaptrick_fft_size(fs, 71.0), K = n / 2 + 1, nap = oracle_num_aperiodicities(fs), D = num_coded_sps;
  const int dim = D + 2 + nap;
  double* x = (double*)malloc(sizeof(double) * xlen);
  for (int i = 0; i < xlen; ++i) x[i] = wave[i] / 32768.0;
  if (preemphasis != 0.0)
    for (int i = xlen - 1; i >= 1; --i) x[i] = x[i] - preemphasis * x[i - 1];
  double* t = (double*)malloc(sizeof(double) * T);
  for (int i = 0; i < T; ++i) t[i] = i * 5.0 / 1000.0;
  double* sp = (double*)malloc(sizeof(double) * (size_t)T * K);
  double* mc = (double*)malloc(sizeof(double) * (size_t)T * D);
  double* coarse = (double*)malloc(sizeof(double) * (size_t)T * (nap > 0 ? nap : 1));
  double* bap = (double*)malloc(sizeof(double) * (size_t)T * (nap > 0 ? nap : 1));
  unsigned char* voiced = (unsigned char*)malloc(T);
  float* lf0 = (float*)malloc(sizeof(float) * T);
  float* vuv = (float*)malloc(sizeof(float) * T);
  oracle_cheaptrick(x, xlen, fs, f0, t, T, n, -0.15, sp);
  int rc = oracle_mcep(sp, 1, T, n, D - 1, alpha, 2, 30, 0.001, 1e-8, mc, NULL);
  oracle_d4c_coarse(x, xlen, fs, f0, t, T, 0.85, coarse, voiced);
  oracle_bap_from_coarse(coarse, voiced, T, fs, n, bap);
  oracle_lf0_vuv(f0, T, 30.0, 0.f, lf0, vuv);
  for (int fr = 0; fr < T; ++fr) {
    float* row = feats + (size_t)fr * dim;
    for (int k = 0; k < D; ++k) row[k] = (float)mc[(size_t)fr * D + k];
    row[D] = lf0[fr];
    row[D + 1] = vuv[fr];
    for (int b = 0; b < nap; ++b) row[D + 2 + b] = (float)bap[(size_t)fr * nap + b];
  }
  free(x); free(t); free(sp); free(mc); free(coarse); free(bap); free(voiced); free(lf0); free(vuv);
  return rc;
}

/* ============================================================================================================================
 * Synthesis half: the per-utterance body of Synthesiser.run_world_synth (idiaptts/src/Synthesiser.py:39-80):
 *   convert_to_world_features (WorldFeatLabelGen.py:735-762) -> decode_sp = exp(Re mgc2sp) as float32
 *   (AudioProcessing.py:248-256, :304-327) -> world_features_to_raw (WorldFeatLabelGen.py:910-945): pyworld.decode_aperiodicity
 *   + pyworld.synthesize.  A C port of oracle/world_np.py::synthesize (WORLD synthesis.cpp restated; PARITY UNPINNED: the
 *   reference holds no golden waveform), pinned against the numpy version by tests/test_oracle_c.py.
 * ========================================================================================================================== */

/* WORLD randn() after randn_reseed(): xorshift128, sum of 12 draws (>> 4), / 2^28 - 6 */
static void xorshift_randn_sequence(double* out, int n) {
  uint32_t x = 123456789u, y = 362436069u, z = 521288629u, w = 88675123u;
  for (int i = 0; i < n; ++i) {
    uint32_t tmp = 0;
    for (int j = 0; j < 12; ++j) {
      const uint32_t t = x ^ (x << 11);
      x = y; y = z; z = w;
      w = (w ^ (w >> 19)) ^ (t ^ (t >> 8));
      tmp += w >> 4;
    }
    out[i] = tmp / 268435456.0 - 6.0;
  }
}

/* WORLD GetMinimumPhaseSpectrum: log-amplitude (K = n/2+1 bins) -> complex minimum-phase spectrum (K bins).
 * wr, wi: scratch of n doubles each. */
static void minimum_phase(const fft_plan* pn, const double* logspec, double* wr, double* wi, double* out_re, double* out_im) {
  const int n = pn->n, h = n / 2;
  for (int i = 0; i <= h; ++i) { wr[i] = logspec[i]; wi[i] = 0.0; }
  for (int i = 1; i < h; ++i) { wr[n - i] = logspec[i]; wi[n - i] = 0.0; }
  cfft(pn, wr, wi);
  /* conj, then fold: [c0, 2 c1 .. 2 c(h-1), c(h), 0 ...] */
  for (int i = 0; i < n; ++i) wi[i] = -wi[i];
  for (int i = 1; i < h; ++i) { wr[i] *= 2.0; wi[i] *= 2.0; }
  for (int i = h + 1; i < n; ++i) { wr[i] = 0.0; wi[i] = 0.0; }
  cfft(pn, wr, wi);
  for (int i = 0; i <= h; ++i) {
    const double a = exp(wr[i] / n), ph = wi[i] / n;
    out_re[i] = a * cos(ph);
    out_im[i] = a * sin(ph);
  }
}

/* unnormalised complex-to-real transform of a half spectrum (numpy irfft * n), result fft-shifted: out[i] = wave[(i + h) % n].
 * The imaginary parts of bins 0 and n/2 are ignored, as numpy's irfft (and FFTW's c2r) do. */
static void c2r_shifted(const fft_plan* pn, const double* xr, const double* xi, double* wr, double* wi, double* out) {
  const int n = pn->n, h = n / 2;
  /* x[m] = sum_k X[k] e^{+2 pi i k m / n} = conj(FFT(conj(X)))[m]; Hermitian extension */
  wr[0] = xr[0]; wi[0] = 0.0;
  wr[h] = xr[h]; wi[h] = 0.0;
  for (int k = 1; k < h; ++k) {
    wr[k] = xr[k]; wi[k] = -xi[k];
    wr[n - k] = xr[k]; wi[n - k] = xi[k];
  }
  cfft(pn, wr, wi);
  for (int i = 0; i < n; ++i) out[i] = wr[(i + h) % n];
}

static double interp1_uniform(double step, const double* y, int len, double xi) {
  /* WORLD interp1 (histc): k with x[k-1] <= xi < x[k], x[i] = i * step, k clipped to [1, len-1] */
  int k0 = (int)(xi / step);
  if (k0 < 0) k0 = 0;
  if (k0 > len - 1) k0 = len - 1;
  while (k0 > 0 && k0 * step > xi) --k0;
  while (k0 + 1 < len && (k0 + 1) * step <= xi) ++k0;
  int k = k0 + 1;
  if (k < 1) k = 1;
  if (k > len - 1) k = len - 1;
  const double s = (xi - (k - 1) * step) / (k * step - (k - 1) * step);
  return y[k - 1] + s * (y[k] - y[k - 1]);
}

/* pyworld.synthesize(f0, sp, ap, fs, frame_period): sp, ap [T x K] row-major; y [int(T * frame_period * fs / 1000)] */
int oracle_synthesize(const double* f0, const double* sp, const double* ap, int T, int fft_size, int fs, double frame_period_ms,
                      double* y, int y_length) {
  const int n = fft_size, h = n / 2, K = h + 1;
  const double fp = frame_period_ms / 1000.0;
  memset(y, 0, sizeof(double) * (size_t)y_length);
  if (T < 2 || y_length < 2) return 0;
  /* ---- time base ---------------------------------------------------------------------------------------------------------- */
  const double lowest_f0 = (double)(fs / n) + 1.0;
  double* cf0 = (double*)malloc(sizeof(double) * (T + 1));
  double* cvuv = (double*)malloc(sizeof(double) * (T + 1));
  for (int i = 0; i < T; ++i) {
    cf0[i] = f0[i] < lowest_f0 ? 0.0 : f0[i];
    cvuv[i] = cf0[i] == 0.0 ? 0.0 : 1.0;
  }
  cf0[T] = cf0[T - 1] * 2 - cf0[T - 2];
  cvuv[T] = cvuv[T - 1] * 2 - cvuv[T - 2];
  double* wrap = (double*)malloc(sizeof(double) * y_length);
  unsigned char* ivuv = (unsigned char*)malloc(y_length);
  const double two_pi = 2.0 * K_PI;
  double total = 0.0;
  for (int i = 0; i < y_length; ++i) {
    const double ti = i / (double)fs;
    double f = interp1_uniform(fp, cf0, T + 1, ti);
    const double v = interp1_uniform(fp, cvuv, T + 1, ti);
    ivuv[i] = v > 0.5 ? 1 : 0;
    if (!ivuv[i]) f = K_DEFAULT_F0;
    total += two_pi * f / fs;
    wrap[i] = fmod(total, two_pi);
  }
  int P = 0;
  int* idx = (int*)malloc(sizeof(int) * y_length);
  double* shift = (double*)malloc(sizeof(double) * y_length);
  for (int i = 0; i + 1 < y_length; ++i) {
    if (fabs(wrap[i + 1] - wrap[i]) > K_PI) {
      const double y1 = wrap[i] - two_pi, y2 = wrap[i + 1];
      idx[P] = i;
      shift[P] = (-y1 / (y2 - y1)) / fs;
      ++P;
    }
  }
  free(cf0); free(cvuv); free(wrap);
  if (P == 0) { free(ivuv); free(idx); free(shift); return 0; }
  /* ---- tables ------------------------------------------------------------------------------------------------------------------ */
  fft_plan* pn = plan_create(n);
  rfft_plan* rp = rplan_create(n);
  double* dc_remover = (double*)malloc(sizeof(double) * n);
  {
    double dc = 0.0;
    for (int i = 0; i < h; ++i) {
      dc_remover[i] = 0.5 - 0.5 * cos(2.0 * K_PI * (i + 1.0) / (1.0 + n));
      dc += dc_remover[i] * 2.0;
    }
    for (int i = 0; i < h; ++i) { dc_remover[i] /= dc; dc_remover[n - 1 - i] = dc_remover[i]; }
  }
  const int total_noise = idx[P - 1] - idx[0];
  double* randn_seq = (double*)malloc(sizeof(double) * (total_noise > 0 ? total_noise : 1));
  xorshift_randn_sequence(randn_seq, total_noise);
  double* env = (double*)malloc(sizeof(double) * K);
  double* ar = (double*)malloc(sizeof(double) * K);
  double* ls = (double*)malloc(sizeof(double) * K);
  double* xr = (double*)malloc(sizeof(double) * K);
  double* xi = (double*)malloc(sizeof(double) * K);
  double* zr = (double*)malloc(sizeof(double) * K);
  double* zi = (double*)malloc(sizeof(double) * K);
  double* wr = (double*)malloc(sizeof(double) * n);
  double* wi = (double*)malloc(sizeof(double) * n);
  double* periodic = (double*)malloc(sizeof(double) * n);
  double* aperiodic = (double*)malloc(sizeof(double) * n);
  double* noise = (double*)malloc(sizeof(double) * n);
  /* ---- pulses ------------------------------------------------------------------------------------------------------------------- */
  for (int p = 0; p < P; ++p) {
    const int n_p = idx[p];
    const int noise_size = idx[imin(P - 1, p + 1)] - n_p;
    const double cur_t = n_p / (double)fs;
    const int fl = imin(T - 1, (int)floor(cur_t / fp)), ce = imin(T - 1, (int)ceil(cur_t / fp));
    const double w = cur_t / fp - fl;
    const double *sp_fl = sp + (size_t)fl * K, *sp_ce = sp + (size_t)ce * K, *ap_fl = ap + (size_t)fl * K, *ap_ce = ap + (size_t)ce * K;
    for (int k = 0; k < K; ++k) {
      double a0 = ap_fl[k] < 0.001 ? 0.001 : (ap_fl[k] > 0.999999999999 ? 0.999999999999 : ap_fl[k]);
      a0 *= a0;
      if (fl == ce) {
        env[k] = fabs(sp_fl[k]);
        ar[k] = a0;
      } else {
        double a1 = ap_ce[k] < 0.001 ? 0.001 : (ap_ce[k] > 0.999999999999 ? 0.999999999999 : ap_ce[k]);
        a1 *= a1;
        env[k] = (1.0 - w) * fabs(sp_fl[k]) + w * fabs(sp_ce[k]);
        ar[k] = (1.0 - w) * a0 + w * a1;
      }
    }
    const int cur_vuv = ivuv[n_p];
    if (!cur_vuv || ar[0] > 0.999) {
      memset(periodic, 0, sizeof(double) * n);
    } else {
      for (int k = 0; k < K; ++k) ls[k] = log(env[k] * (1.0 - ar[k]) + K_SAFE) / 2.0;
      minimum_phase(pn, ls, wr, wi, xr, xi);
      const double coef = 2.0 * K_PI * shift[p] * fs / n;
      for (int k = 0; k < K; ++k) {
        const double re2 = cos(coef * k), im2 = sqrt(1.0 - re2 * re2);
        const double a = xr[k], b = xi[k];
        xr[k] = a * re2 + b * im2;
        xi[k] = b * re2 - a * im2;
      }
      c2r_shifted(pn, xr, xi, wr, wi, periodic);
      double dc = 0.0;
      for (int i = h; i < n; ++i) dc += periodic[i];
      for (int i = 0; i < h; ++i) periodic[i] = -dc * dc_remover[i];
      for (int i = h; i < n; ++i) periodic[i] -= dc * dc_remover[i];
    }
    /* aperiodic response */
    memset(noise, 0, sizeof(double) * n);
    if (noise_size > 0) {
      const double* seg = randn_seq + (n_p - idx[0]);
      double mean = 0.0;
      for (int i = 0; i < noise_size; ++i) mean += seg[i];
      mean /= noise_size;
      for (int i = 0; i < imin(noise_size, n); ++i) noise[i] = seg[i] - mean;
    }
    rfft(rp, noise, zr, zi);
    for (int k = 0; k < K; ++k) ls[k] = (cur_vuv ? log(env[k] * ar[k]) : log(env[k])) / 2.0;
    minimum_phase(pn, ls, wr, wi, xr, xi);
    for (int k = 0; k < K; ++k) {
      const double a = xr[k] * zr[k] - xi[k] * zi[k], b = xr[k] * zi[k] + xi[k] * zr[k];
      xr[k] = a; xi[k] = b;
    }
    c2r_shifted(pn, xr, xi, wr, wi, aperiodic);
    const double sq = sqrt((double)noise_size);
    const int offset = n_p - h + 1;
    const int lo = imax(0, -offset), hi = imin(n, y_length - offset);
    for (int i = lo; i < hi; ++i) y[i + offset] += (periodic[i] * sq + aperiodic[i]) / n;
  }
  free(ivuv); free(idx); free(shift); free(dc_remover); free(randn_seq); free(env); free(ar); free(ls); free(xr); free(xi);
  free(zr); free(zi); free(wr); free(wi); free(periodic); free(aperiodic); free(noise);
  plan_free(pn); rplan_free(rp);
  return 0;
}

/* pyworld.decode_aperiodicity (WORLD codec.cpp): bap [T x nap] -> ap [T x K] */
void oracle_decode_aperiodicity(const double* bap, int T, int fs, int fft_size, double* ap) {
  const int nap = oracle_num_aperiodicities(fs), K = fft_size / 2 + 1;
  for (int fr = 0; fr < T; ++fr) {
    double tmp = 0.0;
    for (int j = 0; j < nap; ++j) tmp += bap[(size_t)fr * nap + j];
    tmp /= nap;
    double* row = ap + (size_t)fr * K;
    for (int k = 0; k < K; ++k)
      row[k] = tmp > -0.5 ? 1.0 - K_SAFE
                          : pow(10.0, coarse_db_at(bap + (size_t)fr * nap, nap, fs / 2.0, (double)k * fs / fft_size) / 20.0);
  }
}

/* One utterance of Synthesiser.run_world_synth: feats [T x (D + 2 + nap)] float32 = [mcep | lf0 | vuv | bap] -> y float32.
 * Returns the number of samples written (int(T * 5 * fs / 1000)), or a negative error code. */
int oracle_synthesize_features(const float* feats, int T, int fs, int num_coded_sps, double alpha, float* y_out, int y_capacity) {
  const int n = oracle_cheaptrick_fft_size(fs, 71.0), h = n / 2, K = h + 1, nap = oracle_num_aperiodicities(fs), D = num_coded_sps;
  const int dim = D + 2 + nap;
  const int y_length = (int)(T * 5.0 * fs / 1000);
  if (y_length > y_capacity) return -1;
  double* f0 = (double*)malloc(sizeof(double) * T);
  double* sp = (double*)malloc(sizeof(double) * (size_t)T * K);
  double* ap = (double*)malloc(sizeof(double) * (size_t)T * K);
  double* bap = (double*)malloc(sizeof(double) * (size_t)T * (nap > 0 ? nap : 1));
  double* c = (double*)malloc(sizeof(double) * (h + 1));
  double* d = (double*)malloc(sizeof(double) * (h + 1));
  double* mc = (double*)malloc(sizeof(double) * D);
  double* buf = (double*)malloc(sizeof(double) * n);
  double* xr = (double*)malloc(sizeof(double) * K);
  double* xi = (double*)malloc(sizeof(double) * K);
  double* y = (double*)malloc(sizeof(double) * (y_length > 0 ? y_length : 1));
  rfft_plan* rp = rplan_create(n);
  for (int fr = 0; fr < T; ++fr) {
    const float* row = feats + (size_t)fr * dim;
    /* world_features_to_raw prologue (W:919-938) */
    double f = exp((double)row[D]);
    int v = row[D + 1] >= 0.5f;
    if (f < 30.0) v = 0;
    f0[fr] = v ? f : 0.0;
    /* decode_sp: pysptk.mgc2sp(mc, alpha, 0, n) = rfft(freqt(mc, n/2, -alpha), n); amp = exp(real) as float32; pow = amp^2 */
    for (int k = 0; k < D; ++k) mc[k] = (double)row[k];
    freqt(mc, D - 1, c, h, -alpha, d);
    memset(buf, 0, sizeof(double) * n);
    for (int k = 0; k <= h; ++k) buf[k] = c[k];
    rfft(rp, buf, xr, xi);
    for (int k = 0; k < K; ++k) {
      const double amp = (double)(float)exp(xr[k]);
      sp[(size_t)fr * K + k] = amp * amp;
    }
    for (int b = 0; b < nap; ++b) bap[(size_t)fr * nap + b] = (double)row[D + 2 + b];
  }
  oracle_decode_aperiodicity(bap, T, fs, n, ap);
  oracle_synthesize(f0, sp, ap, T, n, fs, 5.0, y, y_length);
  for (int i = 0; i < y_length; ++i) y_out[i] = (float)y[i];
  free(f0); free(sp); free(ap); free(bap); free(c); free(d); free(mc); free(buf); free(xr); free(xi); free(y);
  rplan_free(rp);
  return y_length;
}
