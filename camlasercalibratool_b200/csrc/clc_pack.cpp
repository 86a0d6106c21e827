// clc_pack.cpp -- the inner loops of the upload pack threads (host only; compiled by the host compiler, no CUDA).
//
// A run is a contiguous piece of one frame's AoS xyz array (reference Oberserve::points, a std::vector<Eigen::Vector3d>).
// The destination is a pinned DMA slot that the CPU never reads back, so it is written with non-temporal stores: no
// read-for-ownership traffic, no cache pollution -- measured 1.6-2x the rate of the plain loops / memcpy.
// The caller issues clc_pack_fence() once per job before publishing it.
#include <cstdint>
#include <cstring>

#if defined(__x86_64__)
#include <emmintrin.h>
#include <immintrin.h>
#define CLC_X86 1
#else
#define CLC_X86 0
#endif

namespace {

inline uint64_t zbits(const double* z) {
  uint64_t b;
  std::memcpy(&b, z, sizeof(b));
  return b << 1;  // -0.0 is zero; NaN is not
}

#if CLC_X86
// dst is 16-byte aligned (16 bytes per packed point)
uint64_t xy_sse2(const double* src, int64_t n, double* dst) {
  uint64_t any = 0;
  for (int64_t i = 0; i < n; ++i) {
    _mm_stream_pd(dst + 2 * i, _mm_loadu_pd(src + 3 * i));
    any |= zbits(src + 3 * i + 2);
  }
  return any;
}

__attribute__((target("avx2"))) uint64_t xy_avx2(const double* src, int64_t n, double* dst) {
  uint64_t any = 0;
  int64_t i = 0;
  if ((reinterpret_cast<uintptr_t>(dst) & 31u) != 0 && n > 0) {  // one point brings dst to a 32-byte boundary
    _mm_stream_pd(dst, _mm_loadu_pd(src));
    any |= zbits(src + 2);
    i = 1;
  }
  __m256i anyv = _mm256_setzero_si256();
  for (; i + 4 <= n; i += 4) {  // 4 points: 12 doubles in, 8 doubles out
    const __m256d a = _mm256_loadu_pd(src + 3 * i);      // x0 y0 z0 x1
    const __m256d b = _mm256_loadu_pd(src + 3 * i + 4);  // y1 z1 x2 y2
    const __m256d c = _mm256_loadu_pd(src + 3 * i + 8);  // z2 x3 y3 z3
    __m256d o0 = _mm256_blend_pd(a, _mm256_permute4x64_pd(a, 0xFF), 0x4);  // x0 y0 x1 .
    o0 = _mm256_blend_pd(o0, _mm256_permute4x64_pd(b, 0x00), 0x8);        // x0 y0 x1 y1
    const __m256d o1 = _mm256_blend_pd(_mm256_permute4x64_pd(b, 0x0E), _mm256_permute4x64_pd(c, 0x90), 0xC);  // x2 y2 x3 y3
    _mm256_stream_pd(dst + 2 * i, o0);
    _mm256_stream_pd(dst + 2 * i + 4, o1);
    const __m256d z = _mm256_blend_pd(_mm256_blend_pd(a, b, 0x2), c, 0x9);  // z2 z1 z0 z3
    anyv = _mm256_or_si256(anyv, _mm256_slli_epi64(_mm256_castpd_si256(z), 1));
  }
  uint64_t t[4];
  _mm256_storeu_si256(reinterpret_cast<__m256i*>(t), anyv);
  any |= t[0] | t[1] | t[2] | t[3];
  for (; i < n; ++i) {
    _mm_stream_pd(dst + 2 * i, _mm_loadu_pd(src + 3 * i));
    any |= zbits(src + 3 * i + 2);
  }
  return any;
}

// straight copy with non-temporal stores; dst is 8-byte aligned
void copy_nt(const double* src, int64_t n_doubles, double* dst) {
  int64_t i = 0;
  if ((reinterpret_cast<uintptr_t>(dst) & 15u) != 0 && n_doubles > 0) {
    long long v;
    std::memcpy(&v, src, sizeof(v));
    _mm_stream_si64(reinterpret_cast<long long*>(dst), v);
    i = 1;
  }
  for (; i + 2 <= n_doubles; i += 2) _mm_stream_pd(dst + i, _mm_loadu_pd(src + i));
  if (i < n_doubles) {
    long long v;
    std::memcpy(&v, src + i, sizeof(v));
    _mm_stream_si64(reinterpret_cast<long long*>(dst + i), v);
  }
}
#endif

}  // namespace

extern "C" {

// packs x,y of n points (src: AoS xyz) to dst (16-byte aligned); returns non-zero when some z is not exactly zero
uint64_t clc_pack_xy_run(const double* src, int64_t n, double* dst) {
#if CLC_X86
  static const bool avx2 = __builtin_cpu_supports("avx2");
  return avx2 ? xy_avx2(src, n, dst) : xy_sse2(src, n, dst);
#else
  uint64_t any = 0;
  for (int64_t i = 0; i < n; ++i) {
    dst[2 * i] = src[3 * i];
    dst[2 * i + 1] = src[3 * i + 1];
    any |= zbits(src + 3 * i + 2);
  }
  return any;
#endif
}

void clc_pack_xyz_run(const double* src, int64_t n, double* dst) {
#if CLC_X86
  copy_nt(src, 3 * n, dst);
#else
  std::memcpy(dst, src, sizeof(double) * 3 * (size_t)n);
#endif
}

void clc_pack_fence(void) {
#if CLC_X86
  _mm_sfence();
#endif
}

}  // extern "C"
