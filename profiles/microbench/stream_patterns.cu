// stream_patterns.cu -- read-bandwidth microbenchmark behind the design of the sweep kernel (profiles/README.md).
// Reads three SoA double arrays (x, y, z) of N elements once and reduces them, with different work decompositions,
// to find out which access pattern the B200 memory system rewards.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o stream_patterns stream_patterns.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ double2 ldg2(const double* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

__device__ __forceinline__ void block_out(double acc, double* out) {
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// A: classic grid-stride, each thread one double2 per array per iteration (unroll 2)
__global__ void __launch_bounds__(256) k_gridstride(const double* x, const double* y, const double* z, int64_t n, double* out) {
  const int64_t T = (int64_t)gridDim.x * blockDim.x;
  double acc = 0.0;
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  for (; i + 2 * T < n; i += 4 * T) {
    double2 a = ldg2(x + i), b = ldg2(y + i), c = ldg2(z + i);
    double2 d = ldg2(x + i + 2 * T), e = ldg2(y + i + 2 * T), f = ldg2(z + i + 2 * T);
    acc += a.x * b.x + c.x + a.y * b.y + c.y + d.x * e.x + f.x + d.y * e.y + f.y;
  }
  for (; i < n; i += 2 * T) {
    double2 a = ldg2(x + i), b = ldg2(y + i), c = ldg2(z + i);
    acc += a.x * b.x + c.x + a.y * b.y + c.y;
  }
  block_out(acc, out);
}

// B: every warp owns one contiguous range, LDG.128, two 64-point groups in flight
__global__ void __launch_bounds__(256) k_warp_contig_ldg(const double* x, const double* y, const double* z, int64_t n, int64_t per_warp, double* out) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int64_t p0 = gw * per_warp, p1 = p0 + per_warp;
  if (p0 > n) p0 = n;
  if (p1 > n) p1 = n;
  double acc = 0.0;
  int64_t k = p0 + 2 * lane;
  for (; k + 64 < p1; k += 128) {
    double2 a = ldg2(x + k), b = ldg2(y + k), c = ldg2(z + k);
    double2 d = ldg2(x + k + 64), e = ldg2(y + k + 64), f = ldg2(z + k + 64);
    acc += a.x * b.x + c.x + a.y * b.y + c.y + d.x * e.x + f.x + d.y * e.y + f.y;
  }
  if (k < p1) {
    double2 a = ldg2(x + k), b = ldg2(y + k), c = ldg2(z + k);
    acc += a.x * b.x + c.x + a.y * b.y + c.y;
  }
  block_out(acc, out);
}

// C/D/E: per-warp TMA ring.  MODE 0: warp-contiguous ranges; 1: chunks interleaved over all warps of the grid;
// 2: block-contiguous ranges, chunks interleaved over the warps of the block.
template <int CHUNK, int STAGES, int MODE>
__global__ void __launch_bounds__(256) k_tma(const double* x, const double* y, const double* z, int64_t n_chunks_total, int64_t per_warp_chunks,
                                             double* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int64_t gw = (int64_t)blockIdx.x * nw + warp;
  const int64_t W = (int64_t)gridDim.x * nw;
  double* ring = reinterpret_cast<double*>(smem) + (size_t)warp * STAGES * 3 * CHUNK;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)nw * STAGES * 3 * CHUNK * 8) + warp * STAGES;
  // chunk id of this warp's k-th chunk
  auto chunk_id = [&](int64_t k) -> int64_t {
    if (MODE == 0) return gw * per_warp_chunks + k;
    if (MODE == 1) return k * W + gw;
    return ((int64_t)blockIdx.x * per_warp_chunks + k) * nw + warp;  // block range = per_warp_chunks*nw chunks
  };
  int64_t my = 0;
  if (MODE == 0) { int64_t b = gw * per_warp_chunks; my = b >= n_chunks_total ? 0 : (n_chunks_total - b < per_warp_chunks ? n_chunks_total - b : per_warp_chunks); }
  else if (MODE == 1) { my = (n_chunks_total - gw + W - 1) / W; if (gw >= n_chunks_total) my = 0; }
  else { my = 0; for (int64_t k = 0; k < per_warp_chunks; ++k) if (chunk_id(k) < n_chunks_total) my = k + 1; }
  auto issue = [&](int64_t k) {
    const int st = (int)(k % STAGES);
    double* dst = ring + st * 3 * CHUNK;
    const int64_t src = chunk_id(k) * CHUNK;
    mbar_expect_tx(bars + st, 3 * CHUNK * 8);
    bulk_g2s(dst, x + src, CHUNK * 8, bars + st);
    bulk_g2s(dst + CHUNK, y + src, CHUNK * 8, bars + st);
    bulk_g2s(dst + 2 * CHUNK, z + src, CHUNK * 8, bars + st);
  };
  if (lane == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int64_t k = 0; k < STAGES && k < my; ++k) issue(k);
  }
  __syncwarp();
  double acc = 0.0;
  for (int64_t k = 0; k < my; ++k) {
    const int st = (int)(k % STAGES);
    mbar_wait(bars + st, (uint32_t)(k / STAGES) & 1u);
    const double* sx = ring + st * 3 * CHUNK;
#pragma unroll
    for (int g = 0; g < CHUNK / 64; ++g) {
      const double2 a = *reinterpret_cast<const double2*>(sx + g * 64 + 2 * lane);
      const double2 b = *reinterpret_cast<const double2*>(sx + CHUNK + g * 64 + 2 * lane);
      const double2 c = *reinterpret_cast<const double2*>(sx + 2 * CHUNK + g * 64 + 2 * lane);
      acc += a.x * b.x + c.x + a.y * b.y + c.y;
    }
    __syncwarp();
    if (lane == 0 && k + STAGES < my) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(k + STAGES);
    }
  }
  block_out(acc, out);
}

__global__ void k_flush(double* b, int64_t n, double v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) b[i] = v;
}
__global__ void k_fill(double* b, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) b[i] = 1e-3 * (double)(i % 1000);
}

template <typename F>
double time_it(F launch, double* flush, int64_t flush_n, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  std::vector<float> ms;
  for (int r = 0; r < reps + 3; ++r) {
    k_flush<<<592, 256>>>(flush, flush_n, (double)r);
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float m;
    CK(cudaEventElapsedTime(&m, e0, e1));
    if (r >= 3) ms.push_back(m);
  }
  double s = 0;
  for (float m : ms) s += m;
  return s / ms.size();
}

int main(int argc, char** argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : 10000000;  // points
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int64_t n_pad = (n + 4095) / 4096 * 4096 + 4096;
  double *x, *y, *z, *out, *flush;
  CK(cudaMalloc(&x, n_pad * 8)); CK(cudaMalloc(&y, n_pad * 8)); CK(cudaMalloc(&z, n_pad * 8)); CK(cudaMalloc(&out, 8));
  const int64_t flush_n = (256 << 20) / 8;
  CK(cudaMalloc(&flush, flush_n * 8));
  k_fill<<<592, 256>>>(x, n_pad); k_fill<<<592, 256>>>(y, n_pad); k_fill<<<592, 256>>>(z, n_pad);
  CK(cudaDeviceSynchronize());
  const double gb = 24.0 * n / 1e9;
  const int reps = 20;
  printf("n=%lld points, %.1f MB, %d SMs\n", (long long)n, gb * 1e3, sms);
  for (int bps : {2, 4, 8}) {
    const int grid = sms * bps;
    double ms = time_it([&] { k_gridstride<<<grid, 256>>>(x, y, z, n, out); }, flush, flush_n, reps);
    printf("A gridstride_ldg        blocks/SM=%d : %8.2f us  %7.1f GB/s\n", bps, ms * 1e3, gb / (ms * 1e-3));
  }
  for (int bps : {2, 4}) {
    const int grid = sms * bps;
    const int64_t W = (int64_t)grid * 8;
    const int64_t per_warp = ((n + W - 1) / W + 127) / 128 * 128;
    double ms = time_it([&] { k_warp_contig_ldg<<<grid, 256>>>(x, y, z, n, per_warp, out); }, flush, flush_n, reps);
    printf("B warp_contig_ldg       blocks/SM=%d : %8.2f us  %7.1f GB/s\n", bps, ms * 1e3, gb / (ms * 1e-3));
  }
#define RUN_TMA(CHUNK, STAGES, MODE, BPS, NAME)                                                                          \
  {                                                                                                                       \
    const int grid = sms * BPS;                                                                                           \
    const int64_t W = (int64_t)grid * 8;                                                                                  \
    const int64_t nct = (n + CHUNK - 1) / CHUNK;                                                                          \
    const int64_t pwc = (nct + W - 1) / W;                                                                                \
    const size_t smem = (size_t)8 * STAGES * 3 * CHUNK * 8 + 8 * STAGES * 8;                                              \
    CK(cudaFuncSetAttribute(k_tma<CHUNK, STAGES, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
    double ms = time_it([&] { k_tma<CHUNK, STAGES, MODE><<<grid, 256, smem>>>(x, y, z, nct, pwc, out); }, flush, flush_n, reps); \
    printf("%-22s chunk=%4d stages=%d blocks/SM=%d smem=%3zuKB : %8.2f us  %7.1f GB/s\n", NAME, CHUNK, STAGES, BPS,       \
           smem / 1024, ms * 1e3, gb / (ms * 1e-3));                                                                      \
  }
  RUN_TMA(128, 3, 0, 2, "C tma warp-contig");
  RUN_TMA(128, 4, 0, 2, "C tma warp-contig");
  RUN_TMA(256, 3, 0, 2, "C tma warp-contig");
  RUN_TMA(512, 2, 0, 2, "C tma warp-contig");
  RUN_TMA(128, 3, 1, 2, "D tma grid-interleaved");
  RUN_TMA(128, 4, 1, 2, "D tma grid-interleaved");
  RUN_TMA(256, 3, 1, 2, "D tma grid-interleaved");
  RUN_TMA(512, 2, 1, 2, "D tma grid-interleaved");
  RUN_TMA(128, 3, 2, 2, "E tma block-interleaved");
  RUN_TMA(256, 3, 2, 2, "E tma block-interleaved");
  RUN_TMA(128, 6, 1, 1, "D tma grid-interleaved");
  RUN_TMA(256, 4, 1, 1, "D tma grid-interleaved");
  RUN_TMA(128, 6, 0, 1, "C tma warp-contig");
  return 0;
}
