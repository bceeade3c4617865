// Runtime plumbing of libfourierflows_b200: error state, device memory, stream (the B1 seam of SURVEY 8b:
// `zeros(GPU(), T, dims)` src/utils.jl:80, `device_array(GPU())` src/utils.jl:330, upload src/domains.jl:77,
// download `Array(x)` src/output.jl:79).
#include <atomic>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include "ffb_common.cuh"

namespace ffb {

static thread_local char g_err[512] = "";
static cudaStream_t g_stream = nullptr;
static bool g_stream_owned = false;
static std::mutex g_mu;
static std::atomic<uint64_t> g_launches{0};
static int g_num_sms = 0, g_smem_optin = 0;

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

struct ProfRec { std::string name; cudaEvent_t a, b; double bytes; };
static bool g_prof = false;
static std::vector<ProfRec> g_recs;
bool prof_on() { return g_prof; }
void prof_push(const char* name, double bytes) {
  ProfRec r;
  r.name = name; r.bytes = bytes;
  cudaEventCreate(&r.a); cudaEventCreate(&r.b);
  cudaEventRecord(r.a, current_stream());
  g_recs.push_back(r);
}
void prof_pop() { if (!g_recs.empty()) cudaEventRecord(g_recs.back().b, current_stream()); }

cudaStream_t current_stream() {
  if (!g_stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_stream) {
      if (cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking) != cudaSuccess) g_stream = nullptr;
      else g_stream_owned = true;
    }
  }
  return g_stream;
}

static void query_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return;
  cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&g_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
}
int num_sms() { if (!g_num_sms) query_device(); return g_num_sms ? g_num_sms : 148; }
int max_smem_optin() { if (!g_smem_optin) query_device(); return g_smem_optin ? g_smem_optin : 227 * 1024; }

}  // namespace ffb

using namespace ffb;

extern "C" {

const char* ffb_last_error(void) { return g_err; }
int ffb_version(void) { return 100; }

int ffb_device_count(int* n) {
  FFB_REQUIRE(n, FFB_EINVAL, "n is NULL");
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) { *n = 0; return set_error(FFB_ECUDA, "no CUDA device: %s", cudaGetErrorString(e)); }
  *n = c;
  return FFB_OK;
}

int ffb_set_device(int dev) {
  FFB_CUDA(cudaSetDevice(dev));
  g_num_sms = 0; g_smem_optin = 0;
  return FFB_OK;
}

int ffb_set_stream(void* s) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_stream_owned && g_stream) { cudaStreamSynchronize(g_stream); cudaStreamDestroy(g_stream); }
  g_stream = reinterpret_cast<cudaStream_t>(s);
  g_stream_owned = false;
  if (!s) {
    if (cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking) != cudaSuccess)
      return set_error(FFB_ECUDA, "cudaStreamCreate failed");
    g_stream_owned = true;
  }
  return FFB_OK;
}

int ffb_get_stream(void** s) {
  FFB_REQUIRE(s, FFB_EINVAL, "s is NULL");
  *s = current_stream();
  FFB_REQUIRE(*s, FFB_ECUDA, "no CUDA stream (no device?)");
  return FFB_OK;
}

int ffb_sync(void) {
  cudaStream_t st = current_stream();
  FFB_REQUIRE(st, FFB_ECUDA, "no CUDA stream (no device?)");
  FFB_CUDA(cudaStreamSynchronize(st));
  return FFB_OK;
}

int ffb_launch_count(uint64_t* n) {
  FFB_REQUIRE(n, FFB_EINVAL, "n is NULL");
  *n = g_launches.load();
  return FFB_OK;
}

int ffb_prof_enable(int on) {
  for (auto& r : g_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_recs.clear();
  g_prof = on != 0;
  return FFB_OK;
}

// JSON array of {"name", "launches", "ms", "bytes"} aggregated per kernel class since ffb_prof_enable(1)
int ffb_prof_report(char* buf, size_t len) {
  FFB_REQUIRE(buf && len > 2, FFB_EINVAL, "bad buffer");
  cudaStream_t st = current_stream();
  FFB_REQUIRE(st, FFB_ECUDA, "no CUDA stream (no device?)");
  FFB_CUDA(cudaStreamSynchronize(st));
  struct Agg { long n = 0; double ms = 0, bytes = 0; };
  std::map<std::string, Agg> agg;
  for (auto& r : g_recs) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) { cudaGetLastError(); continue; }
    Agg& a = agg[r.name];
    a.n++; a.ms += ms; a.bytes += r.bytes;
  }
  std::string out = "[";
  bool first = true;
  for (auto& kv : agg) {
    char line[320];
    snprintf(line, sizeof(line), "%s{\"name\": \"%s\", \"launches\": %ld, \"ms\": %.6f, \"bytes\": %.0f}", first ? "" : ", ",
             kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.bytes);
    out += line;
    first = false;
  }
  out += "]";
  FFB_REQUIRE(out.size() + 1 <= len, FFB_EINVAL, "report buffer too small (%zu needed)", out.size() + 1);
  memcpy(buf, out.c_str(), out.size() + 1);
  return FFB_OK;
}

int ffb_malloc(void** p, size_t bytes) {
  FFB_REQUIRE(p, FFB_EINVAL, "p is NULL");
  *p = nullptr;
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaMalloc(p, bytes);
  if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return set_error(FFB_ENOMEM, "cudaMalloc(%zu bytes) out of memory", bytes); }
  if (e != cudaSuccess) return set_error(FFB_ECUDA, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
  return FFB_OK;
}

int ffb_free(void* p) {
  if (!p) return FFB_OK;
  // thread-safe (Julia finalizers run on arbitrary threads); cudaFree synchronises the device
  FFB_CUDA(cudaFree(p));
  return FFB_OK;
}

int ffb_memset_zero(void* p, size_t bytes) {
  if (!bytes) return FFB_OK;
  FFB_REQUIRE(p, FFB_EINVAL, "p is NULL");
  FFB_CUDA(cudaMemsetAsync(p, 0, bytes, current_stream()));
  return FFB_OK;
}

int ffb_h2d(void* dst, const void* src, size_t bytes) {
  if (!bytes) return FFB_OK;
  FFB_REQUIRE(dst && src, FFB_EINVAL, "NULL pointer");
  FFB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, current_stream()));
  // pageable sources are staged synchronously by the runtime; pinned ones return immediately
  return FFB_OK;
}

int ffb_d2h(void* dst, const void* src, size_t bytes) {
  if (!bytes) return FFB_OK;
  FFB_REQUIRE(dst && src, FFB_EINVAL, "NULL pointer");
  FFB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, current_stream()));
  FFB_CUDA(cudaStreamSynchronize(current_stream()));
  return FFB_OK;
}

int ffb_d2d(void* dst, const void* src, size_t bytes) {
  if (!bytes) return FFB_OK;
  FFB_REQUIRE(dst && src, FFB_EINVAL, "NULL pointer");
  FFB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, current_stream()));
  return FFB_OK;
}

int ffb_host_alloc_pinned(void** p, size_t bytes) {
  FFB_REQUIRE(p, FFB_EINVAL, "p is NULL");
  FFB_CUDA(cudaMallocHost(p, bytes ? bytes : 16));
  return FFB_OK;
}

int ffb_host_free_pinned(void* p) {
  if (!p) return FFB_OK;
  FFB_CUDA(cudaFreeHost(p));
  return FFB_OK;
}

// ---------------------------------------------------------------- asynchronous output path
// `saveoutput(out)` (src/output.jl:61-79) downloads every field with `Array(data)`, a blocking copy that stalls the step loop.
// A snapshot ring decouples it: ffb_snapshot_begin copies the field into a device staging buffer on the compute stream (HBM
// speed, stream-ordered with the steps), then a dedicated copy stream moves it to pinned host memory while stepping continues;
// ffb_snapshot_wait blocks only the writer.
struct ffb_snapshot {
  size_t bytes; int nbuf, next;
  std::vector<void*> dev, host;
  std::vector<cudaEvent_t> staged, landed;
  std::vector<int> busy;
  cudaStream_t copy_stream;
};

int ffb_snapshot_create(ffb_snapshot** out, size_t bytes, int nbuf) {
  FFB_REQUIRE(out && bytes > 0 && nbuf >= 1 && nbuf <= 64, FFB_EINVAL, "bad argument");
  *out = nullptr;
  auto* s = new ffb_snapshot();
  s->bytes = bytes; s->nbuf = nbuf; s->next = 0; s->copy_stream = nullptr;
  s->dev.assign(nbuf, nullptr); s->host.assign(nbuf, nullptr); s->staged.assign(nbuf, nullptr); s->landed.assign(nbuf, nullptr); s->busy.assign(nbuf, 0);
  *out = s;   // partially built objects are destroyed by the caller through ffb_snapshot_destroy
  FFB_CUDA(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < nbuf; ++i) {
    int rc = ffb_malloc(&s->dev[i], bytes);
    if (rc) return rc;
    FFB_CUDA(cudaMallocHost(&s->host[i], bytes));
    FFB_CUDA(cudaEventCreateWithFlags(&s->staged[i], cudaEventDisableTiming));
    FFB_CUDA(cudaEventCreateWithFlags(&s->landed[i], cudaEventDisableTiming));
  }
  return FFB_OK;
}

int ffb_snapshot_destroy(ffb_snapshot* s) {
  if (!s) return FFB_OK;
  if (s->copy_stream) cudaStreamSynchronize(s->copy_stream);
  for (int i = 0; i < s->nbuf; ++i) {
    if (s->dev[i]) cudaFree(s->dev[i]);
    if (s->host[i]) cudaFreeHost(s->host[i]);
    if (s->staged[i]) cudaEventDestroy(s->staged[i]);
    if (s->landed[i]) cudaEventDestroy(s->landed[i]);
  }
  if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
  delete s;
  return FFB_OK;
}

int ffb_snapshot_begin(ffb_snapshot* s, const void* dev_src, size_t bytes, int* slot) {
  FFB_REQUIRE(s && dev_src && slot, FFB_EINVAL, "NULL argument");
  FFB_REQUIRE(bytes <= s->bytes, FFB_EINVAL, "snapshot of %zu bytes exceeds the ring's %zu", bytes, s->bytes);
  const int i = s->next;
  FFB_REQUIRE(!s->busy[i], FFB_EINVAL, "snapshot ring is full: ffb_snapshot_release slot %d first", i);
  cudaStream_t st = current_stream();
  FFB_REQUIRE(st, FFB_ECUDA, "no CUDA stream (no device?)");
  FFB_CUDA(cudaMemcpyAsync(s->dev[i], dev_src, bytes, cudaMemcpyDeviceToDevice, st));   // ordered with the steps
  FFB_CUDA(cudaEventRecord(s->staged[i], st));
  FFB_CUDA(cudaStreamWaitEvent(s->copy_stream, s->staged[i], 0));
  FFB_CUDA(cudaMemcpyAsync(s->host[i], s->dev[i], bytes, cudaMemcpyDeviceToHost, s->copy_stream));   // beside the next steps
  FFB_CUDA(cudaEventRecord(s->landed[i], s->copy_stream));
  s->busy[i] = 1;
  s->next = (i + 1) % s->nbuf;
  *slot = i;
  return FFB_OK;
}

int ffb_snapshot_wait(ffb_snapshot* s, int slot, void** host_ptr) {
  FFB_REQUIRE(s && host_ptr && slot >= 0 && slot < s->nbuf && s->busy[slot], FFB_EINVAL, "bad slot");
  FFB_CUDA(cudaEventSynchronize(s->landed[slot]));
  *host_ptr = s->host[slot];
  return FFB_OK;
}

int ffb_snapshot_ready(ffb_snapshot* s, int slot, int* ready) {
  FFB_REQUIRE(s && ready && slot >= 0 && slot < s->nbuf && s->busy[slot], FFB_EINVAL, "bad slot");
  cudaError_t e = cudaEventQuery(s->landed[slot]);
  if (e != cudaSuccess && e != cudaErrorNotReady) return set_error(FFB_ECUDA, "cudaEventQuery: %s", cudaGetErrorString(e));
  if (e == cudaErrorNotReady) cudaGetLastError();
  *ready = e == cudaSuccess;
  return FFB_OK;
}

int ffb_snapshot_release(ffb_snapshot* s, int slot) {
  FFB_REQUIRE(s && slot >= 0 && slot < s->nbuf, FFB_EINVAL, "bad slot");
  s->busy[slot] = 0;
  return FFB_OK;
}

int ffb_mem_info(size_t* free_b, size_t* total_b) {
  size_t f = 0, t = 0;
  FFB_CUDA(cudaMemGetInfo(&f, &t));
  if (free_b) *free_b = f;
  if (total_b) *total_b = t;
  return FFB_OK;
}

}  // extern "C"
