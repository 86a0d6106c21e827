// clc_upload.inl -- host memory -> HBM pipeline for the laser points (included by clc_api.cu; host code only).
//
// What the reference's callers hand over is a std::vector<Oberserve>: per frame ONE pageable heap array of Vector3d
// (reference include/LaseCamCalCeres.h:22-23, filled at main/calibr_simulation.cpp:79-103 and
// main/calibr_offline.cpp:144-149).  Copying that with cudaMemcpy frame by frame, or flattening it first on one thread,
// costs several times the PCIe transfer itself.  Here:
//   * a small pool of pack threads gathers the frames into a ring of pinned slots (cudaHostAlloc, kept for the life of
//     the process), one slot = one chunk of `chunk_points` points;
//   * every z is checked on the way (a 2-D laser delivers z == 0 for every point, reference src/utilities.cpp:207): as
//     long as that holds only x,y are packed -- 16 instead of 24 bytes per point cross PCIe -- and the z stream is never
//     created in HBM; the first chunk with a z != 0 is re-packed as xyz and switches the rest of the upload to xyz;
//   * the issuing thread hands every finished slot to the copy engine (its own stream) and queues the AoS -> SoA kernel
//     behind it on the problem's stream, so packing, the PCIe transfer and the layout kernel of consecutive chunks overlap;
//   * several shards (one per device of an in-process multi-GPU group) are fed round-robin from the same pack pool.
// A flat caller buffer that is already pinned (cudaHostAlloc / cudaHostRegister) skips the pack stage.
#pragma once

#include <condition_variable>
#include <functional>
#include <memory>
#include <thread>

namespace {

inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#else
  std::this_thread::yield();
#endif
}

// ---- pack threads: created on first use, parked on a condition variable between uploads ----------------------------
class PackPool {
 public:
  static PackPool& instance() {
    static PackPool pool;
    return pool;
  }
  int size() const { return (int)threads_.size(); }
  // every worker runs fn(worker index) once; returns immediately
  void start(std::function<void(int)> fn) {
    std::lock_guard<std::mutex> lock(m_);
    fn_ = std::move(fn);
    running_ = (int)threads_.size();
    ++generation_;
    cv_work_.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> lock(m_);
    cv_done_.wait(lock, [&] { return running_ == 0; });
  }

 private:
  PackPool() {
    int n = 0;
    if (const char* env = std::getenv("CLC_PACK_THREADS")) n = std::atoi(env);
    if (n <= 0) {
      // Measured (profiles/r2_dropin_*.txt): 8 threads already saturate what the host memory system gives this pipeline, 12 and
      // 16 are no faster.  Stay below the CPUs the process may really use: a container often has a CFS quota far below the
      // visible core count (cgroup v2 cpu.max), and a pool that oversubscribes it gets throttled for whole scheduler periods.
      const int hw = (int)std::thread::hardware_concurrency();
      int usable = hw;
      if (FILE* f = std::fopen("/sys/fs/cgroup/cpu.max", "r")) {
        long long quota = 0, period = 0;
        if (std::fscanf(f, "%lld %lld", &quota, &period) == 2 && quota > 0 && period > 0)
          usable = std::min(usable, (int)(quota / period));
        std::fclose(f);
      }
      n = std::max(1, std::min(16, usable - 1));  // an upload to ONE device uses at most 12 of them (upload_points)
    }
    n = std::min(n, 64);
    for (int i = 0; i < n; ++i) threads_.emplace_back([this, i] { loop(i); });
  }
  ~PackPool() {
    {
      std::lock_guard<std::mutex> lock(m_);
      stop_ = true;
      cv_work_.notify_all();
    }
    for (auto& t : threads_) t.join();
  }
  void loop(int index) {
    uint64_t seen = 0;
    for (;;) {
      std::function<void(int)> fn;
      {
        std::unique_lock<std::mutex> lock(m_);
        cv_work_.wait(lock, [&] { return stop_ || generation_ != seen; });
        if (stop_) return;
        seen = generation_;
        fn = fn_;
      }
      fn(index);
      std::lock_guard<std::mutex> lock(m_);
      if (--running_ == 0) cv_done_.notify_all();
    }
  }
  std::vector<std::thread> threads_;
  std::mutex m_;
  std::condition_variable cv_work_, cv_done_;
  std::function<void(int)> fn_;
  uint64_t generation_ = 0;
  int running_ = 0;
  bool stop_ = false;
};

// ---- pinned slots: portable (usable from every device), grown on demand, never returned -----------------------------
struct PinnedSlots {
  std::vector<void*> slots;
  size_t slot_bytes = 0;
  int ensure(int n, size_t bytes) {
    if (bytes > slot_bytes) {
      for (void* s : slots) cudaFreeHost(s);
      slots.clear();
      slot_bytes = bytes;
    }
    while ((int)slots.size() < n) {
      void* s = nullptr;
      // CLC_UPLOAD_WC=1: write-combined pinned slots (the packers only ever write them, with non-temporal stores, and the DMA
      // engine reads them without snooping the CPU caches)
      unsigned flags = cudaHostAllocPortable;
      if (const char* env = std::getenv("CLC_UPLOAD_WC"))
        if (std::atoi(env) != 0) flags |= cudaHostAllocWriteCombined;
      CLC_CUDA(cudaHostAlloc(&s, slot_bytes, flags));
      slots.push_back(s);
    }
    return CLC_OK;
  }
};
std::mutex g_upload_mutex;  // one upload at a time per process (they would only compete for the same PCIe links and cores)
PinnedSlots g_slots;

// ---- one shard's host data -----------------------------------------------------------------------------------------
struct UploadShard {
  clc_problem* p = nullptr;
  const double* const* frame_points = nullptr;  // gather source: [n_frames] AoS xyz arrays ...
  const double* flat = nullptr;                 // ... or one flat AoS xyz array
  const int64_t* offsets = nullptr;             // [n_frames+1] prefix of the shard (host)
  int64_t n_frames = 0;
  // pipeline state
  cudaStream_t copy_stream = nullptr;
  void* dev_stage[2] = {nullptr, nullptr};
  cudaEvent_t ev_consumed[2] = {nullptr, nullptr};
  int issued = 0;
};

// frame containing local point index q (offsets[f] <= q < offsets[f+1]); frames may be empty
inline int64_t frame_of(const UploadShard& s, int64_t q) {
  return (std::upper_bound(s.offsets, s.offsets + s.n_frames + 1, q) - s.offsets) - 1;
}

template <typename F>
inline void for_each_run(const UploadShard& s, int64_t a, int64_t b, F&& fn) {
  if (a >= b) return;
  if (s.flat) {
    fn(s.flat + 3 * a, b - a);
    return;
  }
  int64_t f = frame_of(s, a), q = a;
  while (q < b) {
    while (s.offsets[f + 1] <= q) ++f;
    const int64_t hi = std::min(b, s.offsets[f + 1]);
    fn(s.frame_points[f] + 3 * (q - s.offsets[f]), hi - q);
    q = hi;
  }
}

// inner loops with non-temporal stores: clc_pack.cpp
extern "C" uint64_t clc_pack_xy_run(const double* src, int64_t n, double* dst);
extern "C" void clc_pack_xyz_run(const double* src, int64_t n, double* dst);
extern "C" void clc_pack_fence(void);

void pack_xyz(const UploadShard& s, int64_t a, int64_t b, double* dst) {
  for_each_run(s, a, b, [&](const double* src, int64_t n) {
    clc_pack_xyz_run(src, n, dst);
    dst += 3 * n;
  });
  clc_pack_fence();
}

// packs x,y only; returns true when some z is not exactly zero (-0.0 counts as zero, NaN does not)
bool pack_xy(const UploadShard& s, int64_t a, int64_t b, double* dst) {
  uint64_t any = 0;
  for_each_run(s, a, b, [&](const double* src, int64_t n) {
    any |= clc_pack_xy_run(src, n, dst);
    dst += 2 * n;
  });
  clc_pack_fence();
  return any != 0;
}

struct UploadStats {  // filled for clc_upload_last_stats (measurement hook)
  double pack_wait_ms = 0.0, total_ms = 0.0, setup_ms = 0.0, issue_ms = 0.0, drain_ms = 0.0;
  int64_t bytes_h2d = 0;
  int chunks = 0, repacked = 0, threads = 0, direct = 0;
};
UploadStats g_last_upload;

// Uploads the points of all shards.  On return every copy and layout kernel has been queued on the shards' streams
// (the caller synchronises), p->z is allocated iff a non-planar chunk was met, and p->z_all_zero / host_planarity_known
// are set (pack path) or p->d_nonplanar will hold the device-side verdict (direct path).
int upload_points(std::vector<UploadShard>& shards) {
  std::lock_guard<std::mutex> upload_lock(g_upload_mutex);
  const auto t_begin = std::chrono::steady_clock::now();
  int64_t chunk_points = (int64_t)256 << 10;
  if (const char* env = std::getenv("CLC_UPLOAD_CHUNK_POINTS")) chunk_points = std::max<int64_t>(1024, std::atoll(env));
  const int G = (int)shards.size();
  int64_t total_points = 0;
  for (auto& s : shards) total_points += s.p->n_points;
  g_last_upload = UploadStats();
  if (total_points == 0) return CLC_OK;

  // direct mode: every shard is one flat, pinned buffer -> no packing, the caller's memory is the DMA source
  bool direct = true;
  for (auto& s : shards) {
    if (!s.flat) { direct = false; break; }
    if (s.p->n_points == 0) continue;
    cudaPointerAttributes attr;
    const cudaError_t e = cudaPointerGetAttributes(&attr, s.flat);
    if (e != cudaSuccess) { cudaGetLastError(); direct = false; break; }
    if (attr.type != cudaMemoryTypeHost) { direct = false; break; }
  }
  if (const char* env = std::getenv("CLC_UPLOAD_DIRECT")) {
    if (std::atoi(env) == 0) direct = false;
  }

  // ---- small uploads (the reference's own sizes: a few thousand points): no pack pool to wake, no extra stream, no events --
  // one pack on the calling thread into a pinned slot, one copy and one layout kernel per shard on the problem's stream.
  if (!direct && total_points <= std::min<int64_t>(65536, chunk_points) && !std::getenv("CLC_UPLOAD_NO_FAST_PATH")) {
    int rc = g_slots.ensure(1, sizeof(double) * 3 * (size_t)chunk_points);
    if (rc != CLC_OK) return rc;
    double* slot = static_cast<double*>(g_slots.slots[0]);
    int64_t bytes_h2d = 0;
    for (auto& s : shards) {
      clc_problem* p = s.p;
      const int64_t n = p->n_points;
      if (n == 0) { p->host_planarity_known = true; p->z_all_zero = true; continue; }
      CLC_CUDA(cudaSetDevice(p->device));
      const bool nonplanar = pack_xy(s, 0, n, slot);
      if (nonplanar) pack_xyz(s, 0, n, slot);
      const size_t bytes = sizeof(double) * (nonplanar ? 3 : 2) * (size_t)n;
      void* stage = nullptr;
      CLC_CUDA(cudaMallocAsync(&stage, bytes, p->stream));
      CLC_CUDA(cudaMemcpyAsync(stage, slot, bytes, cudaMemcpyHostToDevice, p->stream));
      const unsigned blocks = (unsigned)((n + 255) / 256);
      if (nonplanar) {
        int rc2 = materialise_z(p);
        if (rc2 != CLC_OK) return rc2;
        clc::clc_aos_to_soa_kernel<<<blocks, 256, 0, p->stream>>>(static_cast<const double*>(stage), n, p->x, p->y, p->z, 0, p->d_nonplanar);
      } else {
        clc::clc_aos2_to_soa_kernel<<<blocks, 256, 0, p->stream>>>(static_cast<const double2*>(stage), n, p->x, p->y, p->z, 0);
      }
      g_launches.fetch_add(1);
      CLC_CUDA(cudaGetLastError());
      CLC_CUDA(cudaFreeAsync(stage, p->stream));
      CLC_CUDA(cudaStreamSynchronize(p->stream));  // the pinned slot is free again (the next shard / upload packs into it)
      p->host_planarity_known = true;
      p->z_all_zero = !nonplanar;
      bytes_h2d += (int64_t)bytes;
    }
    g_last_upload.bytes_h2d = bytes_h2d;
    g_last_upload.chunks = (int)shards.size();
    g_last_upload.threads = 1;
    g_last_upload.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    return CLC_OK;
  }

  struct Chunk { int shard; int64_t a, b; };
  std::vector<Chunk> chunks;
  {
    std::vector<int64_t> next(G, 0);
    bool more = true;
    while (more) {
      more = false;
      for (int g = 0; g < G; ++g) {
        const int64_t P = shards[g].p->n_points;
        if (next[g] >= P) continue;
        const int64_t b = std::min(P, next[g] + chunk_points);
        chunks.push_back({g, next[g], b});
        next[g] = b;
        more = true;
      }
    }
  }
  const int n_chunks = (int)chunks.size();
  int K = 2 + 2 * G;  // pinned slots in the ring
  if (const char* env = std::getenv("CLC_UPLOAD_SLOTS")) K = std::max(2, std::atoi(env));
  K = std::min(n_chunks, K);
  PackPool* pool = direct ? nullptr : &PackPool::instance();
  // workers used by this upload: one device is saturated by ~8-12 packers (more only oversubscribe a CPU quota and produce
  // outliers); several devices are fed in parallel and are host-bound, so they get the whole pool
  const int parts = direct ? 1 : std::min(pool->size(), 8 + 4 * G);
  if (!direct) {
    int rc = g_slots.ensure(K, sizeof(double) * 3 * (size_t)chunk_points);
    if (rc != CLC_OK) return rc;
  }

  // per-shard device staging (two chunks), copy stream, events
  for (auto& s : shards) {
    clc_problem* p = s.p;
    if (p->n_points == 0) continue;
    CLC_CUDA(cudaSetDevice(p->device));
    CLC_CUDA(cudaStreamCreateWithFlags(&s.copy_stream, cudaStreamNonBlocking));
    const size_t stage_bytes = sizeof(double) * 3 * (size_t)std::min(chunk_points, p->n_points);
    for (int k = 0; k < 2; ++k) {
      CLC_CUDA(cudaMallocAsync(&s.dev_stage[k], stage_bytes, p->stream));
      CLC_CUDA(cudaEventCreateWithFlags(&s.ev_consumed[k], cudaEventDisableTiming));
    }
    // the staging buffers were allocated in stream order on p->stream: the copy stream may use them after this point
    CLC_CUDA(cudaEventRecord(s.ev_consumed[0], p->stream));
    CLC_CUDA(cudaStreamWaitEvent(s.copy_stream, s.ev_consumed[0], 0));
  }

  // ---- pack jobs (workers) ----
  enum : int { kFmtUnset = 0, kFmtXY = 1, kFmtXYZ = 2 };
  std::unique_ptr<std::atomic<int>[]> packed(new std::atomic<int>[n_chunks]);
  std::unique_ptr<std::atomic<int>[]> fmt(new std::atomic<int>[n_chunks]);
  std::unique_ptr<std::atomic<int>[]> bad(new std::atomic<int>[n_chunks]);
  for (int c = 0; c < n_chunks; ++c) { packed[c].store(0); fmt[c].store(direct ? kFmtXYZ : kFmtUnset); bad[c].store(0); }
  std::atomic<int64_t> next_job{0};
  std::atomic<int> allowed{K};  // chunks below this index own a free pinned slot
  std::atomic<bool> nonplanar{false}, abort_flag{false};
  const int64_t n_jobs = (int64_t)n_chunks * parts;
  auto slot_of = [&](int c) { return static_cast<double*>(g_slots.slots[c % K]); };
  auto part_range = [&](const Chunk& ch, int q, int64_t* a, int64_t* b) {
    const int64_t n = ch.b - ch.a;
    *a = ch.a + n * q / parts;
    *b = ch.a + n * (q + 1) / parts;
  };
  if (!direct) {
    pool->start([&](int worker) {
      if (worker >= parts) return;
      for (;;) {
        const int64_t j = next_job.fetch_add(1);
        if (j >= n_jobs) return;
        const int c = (int)(j / parts), q = (int)(j % parts);
        while (c >= allowed.load(std::memory_order_acquire)) {
          if (abort_flag.load()) return;
          cpu_relax();
        }
        int f = fmt[c].load();
        if (f == kFmtUnset) {
          int want = nonplanar.load() ? kFmtXYZ : kFmtXY;
          if (fmt[c].compare_exchange_strong(f, want)) f = want;  // else f holds the winner's choice
        }
        const Chunk& ch = chunks[c];
        int64_t a, b;
        part_range(ch, q, &a, &b);
        double* dst = slot_of(c) + (f == kFmtXY ? 2 : 3) * (a - ch.a);
        if (f == kFmtXY) {
          if (pack_xy(shards[ch.shard], a, b, dst)) {
            bad[c].store(1);
            nonplanar.store(true);
          }
        } else {
          pack_xyz(shards[ch.shard], a, b, dst);
        }
        packed[c].fetch_add(1, std::memory_order_release);
      }
    });
  }

  // ---- issue loop (this thread) ----
  const auto t_setup = std::chrono::steady_clock::now();
  std::vector<cudaEvent_t> ev_copied(n_chunks, nullptr);
  int status = CLC_OK;
  int completed = 0, next_issue = 0;
  double pack_wait_ms = 0.0;
  int repacked = 0;
  int64_t bytes_h2d = 0;
  auto fail_cuda = [&](cudaError_t e, const char* what) {
    status = fail(CLC_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  };
  while (next_issue < n_chunks && status == CLC_OK) {
    // retire finished copies: their pinned slots go back to the packers
    while (completed < next_issue && cudaEventQuery(ev_copied[completed]) == cudaSuccess) {
      ++completed;
      allowed.store(completed + K, std::memory_order_release);
    }
    const int c = next_issue;
    if (!direct && packed[c].load(std::memory_order_acquire) != parts) {
      const auto t0 = std::chrono::steady_clock::now();
      while (packed[c].load(std::memory_order_acquire) != parts) {
        if (completed < next_issue && cudaEventQuery(ev_copied[completed]) == cudaSuccess) {
          ++completed;
          allowed.store(completed + K, std::memory_order_release);
        }
        cpu_relax();
      }
      pack_wait_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    const Chunk& ch = chunks[c];
    UploadShard& s = shards[ch.shard];
    clc_problem* p = s.p;
    const int64_t n = ch.b - ch.a;
    int f = fmt[c].load();
    if (f == kFmtXY && bad[c].load()) {  // a z != 0 turned up in this chunk: it travels as xyz after all
      pack_xyz(s, ch.a, ch.b, slot_of(c));
      f = kFmtXYZ;
      ++repacked;
    }
    cudaError_t e = cudaSetDevice(p->device);
    if (e != cudaSuccess) { fail_cuda(e, "cudaSetDevice"); break; }
    const int ds = s.issued & 1;
    if (s.issued >= 2) e = cudaStreamWaitEvent(s.copy_stream, s.ev_consumed[ds], 0);
    const size_t bytes = sizeof(double) * (f == kFmtXY ? 2 : 3) * (size_t)n;
    const void* src = direct ? static_cast<const void*>(s.flat + 3 * ch.a) : static_cast<const void*>(slot_of(c));
    if (e == cudaSuccess) e = cudaMemcpyAsync(s.dev_stage[ds], src, bytes, cudaMemcpyHostToDevice, s.copy_stream);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_copied[c], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(ev_copied[c], s.copy_stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(p->stream, ev_copied[c], 0);
    if (e != cudaSuccess) { fail_cuda(e, "upload copy"); break; }
    bytes_h2d += (int64_t)bytes;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (f == kFmtXY) {
      clc::clc_aos2_to_soa_kernel<<<blocks, 256, 0, p->stream>>>(static_cast<const double2*>(s.dev_stage[ds]), n, p->x, p->y,
                                                                 p->z, ch.a);
    } else {
      if (p->z == nullptr) {  // first non-planar chunk of this shard: the z stream comes into being, zero so far
        int rc = materialise_z(p);
        if (rc != CLC_OK) { status = rc; break; }
      }
      clc::clc_aos_to_soa_kernel<<<blocks, 256, 0, p->stream>>>(static_cast<const double*>(s.dev_stage[ds]), n, p->x, p->y,
                                                                p->z, ch.a, p->d_nonplanar);
    }
    g_launches.fetch_add(1);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaEventRecord(s.ev_consumed[ds], p->stream);
    if (e != cudaSuccess) { fail_cuda(e, "layout kernel"); break; }
    ++s.issued;
    ++next_issue;
  }
  const auto t_issued = std::chrono::steady_clock::now();
  if (status != CLC_OK) abort_flag.store(true);
  allowed.store(n_chunks + K);  // let the packers run out (they only touch slots nobody reads any more on failure)
  if (!direct) {
    if (status != CLC_OK) next_job.store(n_jobs);
    pool->wait();
  }
  // the pinned slots are reused by the next upload: every copy out of them must be complete before we return
  for (auto& s : shards)
    if (s.copy_stream) {
      cudaSetDevice(s.p->device);
      cudaStreamSynchronize(s.copy_stream);
    }
  for (cudaEvent_t ev : ev_copied)
    if (ev) cudaEventDestroy(ev);
  for (auto& s : shards) {
    if (!s.copy_stream) continue;
    clc_problem* p = s.p;
    cudaSetDevice(p->device);
    for (int k = 0; k < 2; ++k) {
      if (s.dev_stage[k]) cudaFreeAsync(s.dev_stage[k], p->stream);  // stream order: after the last layout kernel
      if (s.ev_consumed[k]) cudaEventDestroy(s.ev_consumed[k]);
    }
    cudaStreamDestroy(s.copy_stream);
    s.copy_stream = nullptr;
    if (!direct) {
      // the packers saw every z: the verdict is known on the host (a shard that never met a non-planar chunk is planar
      // even if another shard of the group was not)
      p->host_planarity_known = true;
      p->z_all_zero = (p->z == nullptr);
    }
  }
  g_last_upload.pack_wait_ms = pack_wait_ms;
  g_last_upload.bytes_h2d = bytes_h2d;
  g_last_upload.chunks = n_chunks;
  g_last_upload.repacked = repacked;
  g_last_upload.threads = direct ? 0 : parts;
  g_last_upload.direct = direct ? 1 : 0;
  const auto t_end = std::chrono::steady_clock::now();
  g_last_upload.total_ms = std::chrono::duration<double, std::milli>(t_end - t_begin).count();
  g_last_upload.setup_ms = std::chrono::duration<double, std::milli>(t_setup - t_begin).count();
  g_last_upload.issue_ms = std::chrono::duration<double, std::milli>(t_issued - t_setup).count();
  g_last_upload.drain_ms = std::chrono::duration<double, std::milli>(t_end - t_issued).count();
  if (std::getenv("CLC_UPLOAD_TIMING"))
    std::fprintf(stderr, "CLC_UPLOAD_TIMING setup_ms=%.3f issue_ms=%.3f (pack_wait_ms=%.3f) drain_ms=%.3f chunks=%d threads=%d MB=%.1f\n",
                 g_last_upload.setup_ms, g_last_upload.issue_ms, pack_wait_ms, g_last_upload.drain_ms, n_chunks, g_last_upload.threads,
                 bytes_h2d / 1e6);
  return status;
}

}  // namespace
