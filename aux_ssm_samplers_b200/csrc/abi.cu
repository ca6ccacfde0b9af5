// libauxssm: error channel, descriptor validation and launch policy shared by all entry points.
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace auxssm {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }

static thread_local char g_log[4096] = "";
static thread_local size_t g_log_len = 0;

void note_dispatch(const char* who) {
  const size_t n = strlen(who);
  if (g_log_len + n + 2 >= sizeof(g_log)) return;  // bounded: later launches are simply not recorded
  if (g_log_len) g_log[g_log_len++] = ';';
  memcpy(g_log + g_log_len, who, n + 1);
  g_log_len += n;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int device_sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  return n;
}

size_t device_max_smem() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 48 * 1024;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || n <= 0)
    return 48 * 1024;
  return (size_t)n;
}

int resolve_time_chunks(const auxssm_lgssm_desc* d) {
  if (!d->parallel || d->T < 3) return 1;
  const int n_elem = d->T - 1;
  if (d->time_chunks > 0) return d->time_chunks < n_elem ? d->time_chunks : n_elem;
  // Automatic policy.  Parallel-in-time costs ~4x the flops of the dense sequential recursion, so it only pays while
  // independent units (chains x inner batch) leave SMs idle: target ~4 CTAs per SM; never more chunks than
  // ~sqrt(6 T) (the cross-chunk prefix pass is sequential).  The STRUCTURED recursion (H = I, diagonal R,
  // time-invariant diagonal F: 2 d^3 flops per step instead of ~14 d^3, against ~45 d^3 per element of the
  // associative formulation) is ~25x cheaper than a scan element, so chunking it only pays for a handful of units.
  const long units = (long)d->n_chains * d->n_batch;
  const uint32_t need = AUXSSM_FLAG_H_IDENTITY | AUXSSM_FLAG_R_DIAGONAL | AUXSSM_FLAG_F_IS_DIAGONAL;
  const bool structured = (d->flags & need) == need && d->Fs.time == 0 && d->Qs.time == 0 && d->bs.time == 0 &&
                          (d->dx > 4);
  const long target = structured ? (device_sm_count() + 15) / 16 : 4L * device_sm_count();
  if (units >= target) return 1;
  long want = (target + units - 1) / units;
  long cap = 1;
  while (cap * cap < 6L * n_elem) ++cap;
  if (want > cap) want = cap;
  if (want > n_elem) want = n_elem;
  return want < 1 ? 1 : (int)want;
}

int check_desc(const auxssm_lgssm_desc* d, const char* who) {
  if (d == nullptr) { set_error("%s: null descriptor", who); return AUXSSM_ERR_INVALID; }
  if (d->T < 1 || d->dx < 1 || d->dy < 1 || d->n_chains < 1 || d->n_batch < 1) {
    set_error("%s: T, dx, dy, n_chains, n_batch must be >= 1 (got %d %d %d %d %d)", who, d->T, d->dx, d->dy,
              d->n_chains, d->n_batch);
    return AUXSSM_ERR_INVALID;
  }
  if (d->dx > AUXSSM_MAX_DIM || d->dy > AUXSSM_MAX_DIM) {
    set_error("%s: dx=%d / dy=%d exceed AUXSSM_MAX_DIM=%d", who, d->dx, d->dy, AUXSSM_MAX_DIM);
    return AUXSSM_ERR_UNSUPPORTED;
  }
  if ((d->flags & AUXSSM_FLAG_H_IDENTITY) && d->dx != d->dy) {
    set_error("%s: AUXSSM_FLAG_H_IDENTITY requires dy == dx", who);
    return AUXSSM_ERR_INVALID;
  }
  if (d->time_chunks < 0) { set_error("%s: time_chunks must be >= 0", who); return AUXSSM_ERR_INVALID; }
  return AUXSSM_OK;
}

}  // namespace auxssm

extern "C" __attribute__((visibility("default"))) const char* auxssm_last_error(void) { return auxssm::g_err; }
extern "C" __attribute__((visibility("default"))) int32_t auxssm_version(void) { return AUXSSM_VERSION; }
extern "C" __attribute__((visibility("default"))) int64_t auxssm_launch_count(void) { return auxssm::launches(); }
extern "C" __attribute__((visibility("default"))) const char* auxssm_dispatch_log(int32_t reset) {
  static thread_local char out[4096];
  memcpy(out, auxssm::g_log, sizeof(out));
  if (reset) { auxssm::g_log[0] = 0; auxssm::g_log_len = 0; }
  return out;
}
