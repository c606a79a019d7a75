// common.cuh — shared definitions of the sm_100a stereo kernels.
//
// Arithmetic contract (DESIGN.md §"Arithmetic"): the reference evaluates everything in IEEE double
// without fused multiply-add (MSVC x64 /O2; the oracle builds with -ffp-contract=off) and the
// refinement is chaotic at the ulp level, so every kernel that produces a value the reference
// would compare or iterate on performs the SAME operations in the SAME order.  All translation
// units are compiled with -fmad=false; the few deliberate fused operations (the exp() twin) use
// __fma_rn explicitly.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SB_NOMATCH (-10000)   // CStereoMatching.h:9
#define SB_MAX_DISPARITY 2    // CStereoMatching.cpp:4
#define SB_MAX_LEVELS 12

struct Bound {  // Boundary, CManageData.h:10-14
  int YL, YR, XL, XR, width, height;
};

// One pyramid level of one view, flat reference layout (pitch = 3*W / W bytes, no padding) so the
// reference's unchecked flat addressing (quirk Q8, Rematch overrun) lands on the same bytes.
struct LevelView {
  const uint8_t* img;   // H*W*3 (+ slack)
  const uint8_t* mask;  // H*W   (+ slack)
};

#define SB_IMG_SLACK 8192  // zero bytes after each image / mask payload

__host__ __device__ inline int sb_imin(int a, int b) { return a > b ? b : a; }
__host__ __device__ inline int sb_imax(int a, int b) { return a < b ? b : a; }

// ------------------------------------------------------------------------------------------------
// exp(): bit-for-bit twin of the C library the oracle links (glibc 2.39 x86-64, FMA variant).
// Algorithm as published with the library (table-driven, N = 128, degree-5 polynomial); the
// placement of the fused operations follows the library build.  tests/test_exp_twin.py pins the
// host twin against libm on the test machine.  Domain used by the path: x = -(e*e) <= 0.
// ------------------------------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
#define SB_FMA(a, b, c) __fma_rn((a), (b), (c))
#define SB_ASU(x) ((unsigned long long)__double_as_longlong(x))
#define SB_ASD(u) __longlong_as_double((long long)(u))
#else
#include <cmath>
#include <cstring>
#define SB_FMA(a, b, c) std::fma((a), (b), (c))
static inline unsigned long long sb_asu_host(double x) { unsigned long long u; memcpy(&u, &x, 8); return u; }
static inline double sb_asd_host(unsigned long long u) { double x; memcpy(&x, &u, 8); return x; }
#define SB_ASU(x) sb_asu_host(x)
#define SB_ASD(u) sb_asd_host(u)
#endif

__host__ __device__ inline double sb_exp_twin(double x, const unsigned long long* __restrict__ tab) {
  const double InvLn2N = 0x1.71547652b82fep+7, Shift = 0x1.8p52;
  const double NegLn2hiN = -0x1.62e42fefa0000p-8, NegLn2loN = -0x1.cf79abc9e3b3ap-47;
  const double C2 = 0x1.ffffffffffdbdp-2, C3 = 0x1.555555555543cp-3, C4 = 0x1.55555cf172b91p-5,
               C5 = 0x1.1111167a4d017p-7;
  const unsigned long long bx = SB_ASU(x);
  unsigned abstop = (unsigned)(bx >> 52) & 0x7ffu;
  if (abstop - 0x3c9u > 0x3eu) {
    if ((int)(abstop - 0x3c9u) < 0) return 1.0 + x;  // |x| < 2^-54
    if (abstop > 0x408u) {                           // |x| >= 1024, inf, nan
      if (bx == 0xfff0000000000000ull) return 0.0;
      if (abstop == 0x7ffu) return 1.0 + x;
      if (bx >> 63) return 0.0;  // underflow
#ifdef __CUDA_ARCH__
      return __longlong_as_double(0x7ff0000000000000ll);
#else
      return INFINITY;
#endif
    }
    abstop = 0;  // 512 <= |x| < 1024: handled after the polynomial
  }
  double kd = SB_FMA(x, InvLn2N, Shift);
  const unsigned long long ki = SB_ASU(kd);
  kd = kd - Shift;
  const double r = SB_FMA(kd, NegLn2loN, SB_FMA(kd, NegLn2hiN, x));
  const unsigned idx = 2u * (unsigned)(ki & 127u);
  unsigned long long sbits = tab[idx + 1] + (ki << 45);
  const double tail = SB_ASD(tab[idx]);
  const double p23 = SB_FMA(r, C3, C2);
  const double tr = r + tail;
  const double r2 = r * r;
  const double p45 = SB_FMA(r, C5, C4);
  const double t1 = SB_FMA(p23, r2, tr);
  const double r4 = r2 * r2;
  const double tmp = SB_FMA(r4, p45, t1);
  if (abstop == 0) {
    if ((ki & 0x80000000ull) == 0) {  // k > 0: x >= 512, never reached with x <= 0
      sbits -= 1009ull << 52;
      const double s = SB_ASD(sbits);
      return 0x1p1009 * SB_FMA(s, tmp, s);
    }
    sbits += 1022ull << 52;
    const double scale = SB_ASD(sbits);
    const double st = scale * tmp;
    double y = scale + st;
    if (y < 1.0) {
      const double hi = 1.0 + y;
      double lo = scale - y;
      lo = lo + st;
      double t = 1.0 - hi;
      t = t + y;
      t = t + lo;
      y = (t + hi) - 1.0;
      if (y == 0.0) y = 0.0;
    }
    return 0x1p-1022 * y;
  }
  const double scale = SB_ASD(sbits);
  return SB_FMA(scale, tmp, scale);
}

#ifdef __CUDACC__
// u8 -> f64, exact.
__device__ __forceinline__ double sb_u2d(unsigned v) { return (double)v; }

// NWORDS 32-bit words starting at byte offset bo of a byte buffer (any alignment): aligned word loads + funnel shifts.
// Reads up to 4 * (NWORDS + 1) bytes from the aligned address below bo; the image buffers carry slack for that.
template <int NWORDS>
__device__ __forceinline__ void load_row_words(const uint8_t* __restrict__ base, long bo, unsigned (&w)[NWORDS]) {
  const unsigned sh = ((unsigned)bo & 3u) * 8u;
  const unsigned* __restrict__ p = reinterpret_cast<const unsigned*>(base + (bo & ~3L));  
  unsigned a[NWORDS + 1];
#pragma unroll
  for (int i = 0; i <= NWORDS; i++) a[i] = p[i];
#pragma unroll
  for (int i = 0; i < NWORDS; i++) w[i] = __funnelshift_r(a[i], a[i + 1], sh);
}

// First-max argmax merge used by every NCC search (strict '>' in ascending candidate order,
// CStereoMatching.cpp:214-218): (v, i) beats (bv, bi) iff v > bv, or v == bv and i < bi.
__device__ __forceinline__ void sb_argmax_merge(double& bv, int& bi, double v, int i) {
  if (v > bv || (v == bv && i < bi && i >= 0)) { bv = v; bi = i; }
}
#endif
