// Error string + device-property cache behind the C-ABI.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>

namespace nabu {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const char* last_error() { return g_err; }

void warn_once(const char* key, const char* fmt, ...) {
  static char seen[64][96];
  static int nseen = 0;
  static int quiet = -1;
  if (quiet < 0) quiet = (getenv("NABU_QUIET") && atoi(getenv("NABU_QUIET")) != 0) ? 1 : 0;
  if (quiet) return;
  for (int i = 0; i < nseen; ++i)
    if (strncmp(seen[i], key, sizeof(seen[i]) - 1) == 0) return;
  if (nseen < 64) {
    strncpy(seen[nseen], key, sizeof(seen[nseen]) - 1);
    seen[nseen][sizeof(seen[nseen]) - 1] = 0;
    ++nseen;
  }
  va_list ap;
  va_start(ap, fmt);
  fprintf(stderr, "[nabu_b200] ");
  vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n");
  va_end(ap);
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached[dev] = v;
  }
  return cached[dev];
}

int max_smem_optin() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 227 * 1024;
  if (cached[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || v <= 0)
      v = 227 * 1024;
    cached[dev] = v;
  }
  return cached[dev];
}


// ---- deferred weight gradients ---------------------------------------------------------------------
Overlap& overlap() {
  static Overlap o = {};
  return o;
}
int overlap_init() {
  Overlap& o = overlap();
  if (o.hp) return 0;
  int lo = 0, hi = 0;
  NABU_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  NABU_CHECK_CUDA(cudaStreamCreateWithPriority(&o.hp, cudaStreamNonBlocking, hi));
  NABU_CHECK_CUDA(cudaStreamCreateWithPriority(&o.side, cudaStreamNonBlocking, lo));
  NABU_CHECK_CUDA(cudaEventCreateWithFlags(&o.ev_pre, cudaEventDisableTiming));
  NABU_CHECK_CUDA(cudaEventCreateWithFlags(&o.ev_rec, cudaEventDisableTiming));
  NABU_CHECK_CUDA(cudaEventCreateWithFlags(&o.ev_done, cudaEventDisableTiming));
  return 0;
}
int overlap_workspace(size_t bytes) {
  Overlap& o = overlap();
  if (bytes <= o.ws_bytes) return 0;
  if (o.ws) {
    NABU_CHECK_CUDA(cudaDeviceSynchronize());          // the caller's stream may still read operand planes that live here
    NABU_CHECK_CUDA(cudaFree(o.ws));
    o.ws = nullptr; o.ws_bytes = 0;
  }
  NABU_CHECK_CUDA(cudaMalloc(&o.ws, bytes));
  o.ws_bytes = bytes;
  return 0;
}
int overlap_join(cudaStream_t stream) {
  Overlap& o = overlap();
  if (!o.pending) return 0;
  NABU_CHECK_CUDA(cudaStreamWaitEvent(stream, o.ev_done, 0));
  o.pending = false;
  return 0;
}

// ---- launch counter + optional per-kernel event timing ----------------------------------------
namespace {
struct ProfRec { const char* name; cudaEvent_t a, b; };
constexpr int kMaxRec = 1 << 16;
ProfRec g_rec[kMaxRec];
int g_nrec = 0;
bool g_prof = false;
unsigned long long g_launches = 0;
}  // namespace

KernelScope::KernelScope(const char* name, cudaStream_t stream) : stream_(stream), slot_(-1) {
  ++g_launches;
  if (g_prof && g_nrec < kMaxRec) {
    slot_ = g_nrec++;
    ProfRec& r = g_rec[slot_];
    r.name = name;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, stream);
  }
}
KernelScope::~KernelScope() {
  if (slot_ >= 0) cudaEventRecord(g_rec[slot_].b, stream_);
}

unsigned long long kernel_launches() { return g_launches; }
void profile_enable(bool on) { g_prof = on; }

// Synchronises, then writes {"name": [count, total_ms], ...} and frees the events.
int profile_collect(char* out, size_t cap) {
  cudaDeviceSynchronize();
  struct Agg { const char* name; int n; double ms; };
  Agg agg[256];
  int na = 0;
  for (int i = 0; i < g_nrec; ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, g_rec[i].a, g_rec[i].b);
    cudaEventDestroy(g_rec[i].a);
    cudaEventDestroy(g_rec[i].b);
    int j = 0;
    for (; j < na; ++j)
      if (strcmp(agg[j].name, g_rec[i].name) == 0) break;
    if (j == na) {
      if (na == 256) continue;
      agg[na].name = g_rec[i].name; agg[na].n = 0; agg[na].ms = 0.0; ++na;
    }
    agg[j].n += 1;
    agg[j].ms += ms;
  }
  g_nrec = 0;
  size_t off = 0;
  auto put = [&](const char* fmt, auto... a) {
    if (off < cap) off += snprintf(out + off, cap - off, fmt, a...);
  };
  put("{");
  for (int j = 0; j < na; ++j) put("%s\"%s\": [%d, %.6f]", j ? ", " : "", agg[j].name, agg[j].n, agg[j].ms);
  put("}");
  return off < cap ? 0 : 1;
}

}  // namespace nabu
