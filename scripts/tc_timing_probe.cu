// In-kernel timeline of the tcgen05 kernels, through the C ABI (include/dfb200.h). Built by scripts/build_timing_probe.sh
// with -DDFB_TC_TIMING (the shipped library has no stamps). For each problem: a few warm-up calls, then one call whose
// first and last CTA recorded clock64() at the hand-over points; printed as microseconds since the CTA's start.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../include/dfb200.h"

extern "C" int dfb_debug_tc_stamps(unsigned long long* out32);

static const char* kNames[14] = {"entry", "setup done", "pdl_sync done", "first TMA issued", "last TMA issued", "first stage full",
                                 "last MMA committed", "accumulator ready", "tile out of TMEM", "finish()", "cluster barrier 1",
                                 "cluster reduce done", "final barrier", "TMEM freed"};
static double g_mhz = 1900.0;

static void report(const char* what, float ms_per_call) {
  unsigned long long st[32];
  dfb_synchronize();
  if (dfb_debug_tc_stamps(st) != 0) { printf("stamps unavailable\n"); return; }
  printf("\n%s : %.2f us per call (events around 20 back-to-back calls)\n", what, ms_per_call * 1e3);
  for (int slot = 0; slot < 2; ++slot) {
    printf("  %s CTA:", slot == 0 ? "first" : "last ");
    const unsigned long long t0 = st[slot * 16];
    for (int i = 1; i < 14; ++i) {
      const unsigned long long t = st[slot * 16 + i];
      if (t >= t0 && t - t0 < 100000000ull) printf("  %s %.2f", kNames[i], (double)(t - t0) / g_mhz);
    }
    printf("\n");
  }
}

template <class F>
static float time_calls(F f) {
  for (int i = 0; i < 5; ++i) f();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  dfb_synchronize();
  // the library's compute stream is not visible here: bracket with device-wide synchronisation and host timing of
  // 20 calls is too coarse, so use events on the legacy stream after a full synchronize (the calls are stream ordered)
  cudaEventRecord(e0, 0);
  for (int i = 0; i < 20; ++i) f();
  dfb_synchronize();
  cudaEventRecord(e1, 0);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / 20.f;
}

int main() {
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  if (khz > 0) g_mhz = khz / 1000.0;
  printf("SM clock (max) %.0f MHz; stamps are SM cycles / that\n", g_mhz);
  auto alloc = [](size_t n, float v) { float* p = nullptr; if (dfb_malloc(n, &p) != DFB_OK) { printf("alloc failed: %s\n", dfb_last_error()); exit(1); } dfb_fill(p, v, n); return p; };
  struct G { int M, N, K; };
  for (G g : {G{1024, 256, 32}, G{1024, 256, 576}, G{1024, 256, 2304}, G{16384, 64, 32}, G{65536, 32, 288}}) {
    float *A = alloc((size_t)g.M * g.K, 0.5f), *B = alloc((size_t)g.N * g.K, 0.25f), *C = alloc((size_t)g.M * g.N, 0.f);
    auto f = [&] { if (dfb_gemm(A, B, C, g.M, g.N, g.K, 0, 1, g.K, g.K, g.N, 0, nullptr, DFB_MODE_TF32) != DFB_OK) { printf("gemm: %s\n", dfb_last_error()); exit(1); } };
    char name[128];
    snprintf(name, sizeof name, "gemm M=%d N=%d K=%d (B K-major)", g.M, g.N, g.K);
    float ms = time_calls(f);
    report(name, ms);
    dfb_free(A); dfb_free(B); dfb_free(C);
  }
  struct Cv { int n, c, h, k, s; };
  for (Cv v : {Cv{256, 32, 16, 32, 1}, Cv{256, 64, 8, 64, 1}, Cv{256, 128, 4, 128, 1}, Cv{256, 256, 2, 256, 1}}) {
    const int oh = (v.h + 2 - 3) / v.s + 1;
    float *x = alloc((size_t)v.n * v.c * v.h * v.h, 0.5f), *w = alloc((size_t)v.k * v.c * 9, 0.1f), *y = alloc((size_t)v.n * oh * oh * v.k, 0.f);
    float *dx = alloc((size_t)v.n * v.c * v.h * v.h, 0.f), *dw = alloc((size_t)v.k * v.c * 9, 0.f);
    char name[160];
    auto fp = [&] { if (dfb_conv2d_fprop(x, DFB_LAYOUT_NHWC, w, DFB_WLAYOUT_KRSC, y, v.n, v.c, v.h, v.h, v.k, 3, 1, v.s, DFB_MODE_TF32, nullptr, 0) != DFB_OK) { printf("fprop: %s\n", dfb_last_error()); exit(1); } };
    snprintf(name, sizeof name, "conv fprop %dx%dx%dx%d -> %d", v.n, v.c, v.h, v.h, v.k);
    float ms = time_calls(fp);
    report(name, ms);
    auto dg = [&] { if (dfb_conv2d_dgrad(y, w, DFB_WLAYOUT_KRSC, dx, v.n, v.c, v.h, v.h, v.k, 3, 1, v.s, DFB_MODE_TF32, DFB_DGRAD_EXACT, nullptr, 0) != DFB_OK) { printf("dgrad: %s\n", dfb_last_error()); exit(1); } };
    snprintf(name, sizeof name, "conv dgrad %dx%dx%dx%d -> %d", v.n, v.c, v.h, v.h, v.k);
    ms = time_calls(dg);
    report(name, ms);
    auto wg = [&] { if (dfb_conv2d_wgrad(x, DFB_LAYOUT_NHWC, y, dw, DFB_WLAYOUT_KRSC, v.n, v.c, v.h, v.h, v.k, 3, 1, v.s, DFB_MODE_TF32, nullptr, 0) != DFB_OK) { printf("wgrad: %s\n", dfb_last_error()); exit(1); } };
    snprintf(name, sizeof name, "conv wgrad %dx%dx%dx%d -> %d", v.n, v.c, v.h, v.h, v.k);
    ms = time_calls(wg);
    report(name, ms);
    dfb_free(x); dfb_free(w); dfb_free(y); dfb_free(dx); dfb_free(dw);
  }
  return 0;
}
