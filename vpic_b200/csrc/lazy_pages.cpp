// Page-protected lazy coherence: see lazy_pages.h for the contract.
#include "lazy_pages.h"

#include <atomic>
#include <errno.h>
#include <fcntl.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

namespace vpb_lazy {

namespace {

enum : uint8_t { HOST = 0, DEVICE = 1 };

constexpr int kMaxRegions = 512;
constexpr size_t kGuard = 64;

}  // namespace

struct Region {
  char *h = nullptr; size_t cap = 0; char *d = nullptr;
  char *lo = nullptr, *hi = nullptr;     // whole pages inside [h, h+cap)
  char *base = nullptr;                  // chunk 0 starts here (<= lo, on a chunk boundary)
  size_t nchunks = 0, ndevice = 0;
  uint8_t *state = nullptr;
  size_t last_fault_end = (size_t)-1;    // chunk after the last window fetched by the fault handler
  size_t window = 1;
  // The two unprotected ends are copied both ways on every call.  shadow holds what host and device agreed on at the last
  // of those copies (head edge first, then the tail edge); the leading head_ok / tail_ok bytes of each are valid.  An
  // end the host has not changed since (memcmp) is not uploaded again.
  char *shadow = nullptr;
  size_t head_ok = 0, tail_ok = 0;
  bool refresh_head = false, refresh_tail = false;   // shadow to be re-read from the host once the queued copy-back is done
  size_t refresh_head_len = 0, refresh_tail_len = 0;
};

namespace {

Copier g_copy;
size_t g_chunk = 2u << 20;
size_t g_page = 4096;
Region *g_regions[kMaxRegions];
std::atomic<int> g_nregions{0};
std::atomic<long> g_owner{0};            // tid of the thread inside the tracker, 0 = free
struct sigaction g_old_segv;
bool g_installed = false;
int g_probe[2] = {-1, -1};           // pipe: write(2) from a PROT_NONE page fails with EFAULT
Stats g_stats = {0, 0, 0, 0};
int g_mem_fd = -1;                       // /proc/self/mem: writes through it ignore page protection (FOLL_FORCE)
char *g_staging = nullptr;
constexpr size_t kStaging = 8u << 20;

Region *g_refresh[kMaxRegions];          // regions whose shadow waits for an asynchronous copy-back (after_sync)
int g_nrefresh = 0;

long self_tid() { return (long)syscall(SYS_gettid); }

struct Lock {
  bool reentered = false;
  Lock() {
    const long me = self_tid();
    if (g_owner.load(std::memory_order_relaxed) == me) { reentered = true; return; }
    long expect = 0;
    while (!g_owner.compare_exchange_weak(expect, me, std::memory_order_acquire)) { expect = 0; }
  }
  ~Lock() { if (!reentered) g_owner.store(0, std::memory_order_release); }
};

void fatal(const char *msg) {
  if (g_copy.fatal) g_copy.fatal(msg);
  (void)!syscall(SYS_write, 2, msg, strlen(msg)); (void)!syscall(SYS_write, 2, "\n", 1);
  _exit(1);
}

inline char *chunk_lo(const Region *r, size_t c) { char *p = r->base + c * g_chunk; return p < r->lo ? r->lo : p; }
inline char *chunk_hi(const Region *r, size_t c) { char *p = r->base + (c + 1) * g_chunk; return p > r->hi ? r->hi : p; }
inline size_t chunk_of(const Region *r, const char *p) { return (size_t)(p - r->base) / g_chunk; }

void protect(const Region *r, size_t c0, size_t c1, int prot) {      // chunks [c0, c1)
  if (c0 >= c1) return;
  char *a = chunk_lo(r, c0), *b = chunk_hi(r, c1 - 1);
  if (mprotect(a, (size_t)(b - a), prot) != 0 && prot == PROT_NONE) fatal("vpic_b200: mprotect failed on a host array");
}

// Fill still-protected host pages [a, a+n) from the device.  False when neither protection-blind route works.
bool fill_protected(char *a, const char *dev, size_t n) {
  if (g_copy.d2h_protected && g_copy.d2h_protected(a, dev, n) == 0) return true;
  if (g_mem_fd < 0 || !g_staging) return false;          // the staging buffer is allocated in init(), never in here
  for (size_t off = 0; off < n; off += kStaging) {
    const size_t m = n - off < kStaging ? n - off : kStaging;
    if (g_copy.d2h(g_staging, dev + off, m)) fatal("vpic_b200: device-to-host copy failed while serving a host access");
    size_t done = 0;
    while (done < m) {
      // raw system call: the library interposes pwrite(2) for the host program (dropin.cu)
      const ssize_t w = syscall(SYS_pwrite64, g_mem_fd, g_staging + done, m - done, (off_t)(uintptr_t)(a + off + done));
      if (w <= 0) {
        if (off == 0 && done == 0) { close(g_mem_fd); g_mem_fd = -1; return false; }   // not permitted here: fall back
        fatal("vpic_b200: write through /proc/self/mem failed half way");
      }
      done += (size_t)w;
    }
  }
  return true;
}

// device-owned chunks [c0, c1) -> host-owned: copy back, then unprotect
void fetch(Region *r, size_t c0, size_t c1, uint64_t *bytes) {
  size_t c = c0;
  while (c < c1) {
    if (r->state[c] != DEVICE) { c++; continue; }
    size_t e = c;
    while (e < c1 && r->state[e] == DEVICE) e++;
    char *a = chunk_lo(r, c), *b = chunk_hi(r, e - 1);
    if (fill_protected(a, r->d + (a - r->h), (size_t)(b - a))) {
      protect(r, c, e, PROT_READ | PROT_WRITE);
    } else {
      // last resort (no page-locked memory, no /proc/self/mem): open the pages first.  Correct for a host that
      // touches its arrays from one thread at a time; a second thread could see the chunk half filled.
      protect(r, c, e, PROT_READ | PROT_WRITE);
      if (g_copy.d2h(a, r->d + (a - r->h), (size_t)(b - a))) fatal("vpic_b200: device-to-host copy failed while serving a host access");
    }
    if (bytes) *bytes += (uint64_t)(b - a);
    for (size_t k = c; k < e; k++) r->state[k] = HOST;
    r->ndevice -= e - c;
    c = e;
  }
}

// A device-owned chunk must still be inaccessible.  If the host freed the array and the allocator mapped fresh
// memory at the same address the protection is gone: the device copy describes memory that no longer exists.
bool still_protected(const Region *r) {
  for (size_t c = 0; c < r->nchunks; c++) {
    if (r->state[c] != DEVICE) continue;
    if (g_probe[1] < 0) return true;
    // raw system calls: write(2)/read(2) themselves are interposed for the host program (dropin.cu) and would fetch
    const ssize_t n = syscall(SYS_write, g_probe[1], chunk_lo(r, c), 1);   // the kernel refuses to read PROT_NONE memory
    if (n == 1) { char b; (void)!syscall(SYS_read, g_probe[0], &b, 1); return false; }
    return errno == EFAULT;
  }
  return true;
}

Region *find(const char *p) {
  const int n = g_nregions.load(std::memory_order_acquire);
  for (int i = 0; i < n; i++) { Region *r = g_regions[i]; if (r && p >= r->lo && p < r->hi) return r; }
  return nullptr;
}

void chain(int sig, siginfo_t *si, void *uc) {
  if (g_old_segv.sa_flags & SA_SIGINFO) {
    if (g_old_segv.sa_sigaction) { g_old_segv.sa_sigaction(sig, si, uc); return; }
  } else if (g_old_segv.sa_handler != SIG_DFL && g_old_segv.sa_handler != SIG_IGN) {
    g_old_segv.sa_handler(sig); return;
  }
  // default action: put it back and let the access fault again
  struct sigaction dfl; memset(&dfl, 0, sizeof dfl); dfl.sa_handler = SIG_DFL; sigemptyset(&dfl.sa_mask);
  sigaction(SIGSEGV, &dfl, nullptr);
}

void on_segv(int sig, siginfo_t *si, void *uc) {
  const int saved_errno = errno;
  const char *addr = (const char *)si->si_addr;
  if (si->si_code != SEGV_ACCERR || g_owner.load(std::memory_order_relaxed) == self_tid()) { chain(sig, si, uc); return; }
  bool ours = false;
  {
    Lock lk;
    Region *r = find(addr);
    if (r) {
      ours = true;
      const size_t c = chunk_of(r, addr);
      if (r->state[c] == DEVICE) {
        // sequential readers (dumps, accumulate_hydro_p) get a growing window per fault
        r->window = (c == r->last_fault_end) ? (r->window < 32 ? r->window * 2 : 32) : 1;
        size_t e = c + r->window; if (e > r->nchunks) e = r->nchunks;
        uint64_t b = 0;
        fetch(r, c, e, &b);
        r->last_fault_end = e;
        g_stats.faults++; g_stats.fault_bytes += b;
      }   // else: another thread fetched it first; just retry
    }
  }
  if (!ours) chain(sig, si, uc);
  errno = saved_errno;
}

void install_handler() {
  struct sigaction sa; memset(&sa, 0, sizeof sa);
  sa.sa_sigaction = on_segv; sa.sa_flags = SA_SIGINFO | SA_NODEFER;
  sigemptyset(&sa.sa_mask);
  struct sigaction prev;
  if (sigaction(SIGSEGV, &sa, &prev) != 0) fatal("vpic_b200: cannot install the SIGSEGV handler");
  if (!((prev.sa_flags & SA_SIGINFO) && prev.sa_sigaction == on_segv)) g_old_segv = prev;   // chain to whoever was there
}

// The host program (an MPI runtime, a crash reporter) may install its own SIGSEGV handler after ours.  Pages are only
// protected from to_device(), so checking there is enough: take the signal back and chain to the newcomer.
void ensure_handler() {
  struct sigaction cur;
  if (sigaction(SIGSEGV, nullptr, &cur) == 0 && (cur.sa_flags & SA_SIGINFO) && cur.sa_sigaction == on_segv) return;
  install_handler();
}

void at_exit() {
  // hand everything back before the process tears down (the host's destructors may walk its arrays)
  Lock lk;
  const int n = g_nregions.load();
  for (int i = 0; i < n; i++) {
    Region *r = g_regions[i];
    if (r && r->ndevice) mprotect(r->lo, (size_t)(r->hi - r->lo), PROT_READ | PROT_WRITE);
  }
}

}  // namespace

void init(const Copier &c, size_t chunk_bytes) {
  Lock lk;
  g_copy = c;
  if (g_installed) return;
  g_page = (size_t)sysconf(_SC_PAGESIZE);
  if (chunk_bytes) g_chunk = chunk_bytes < g_page ? g_page : (chunk_bytes / g_page) * g_page;
  if (pipe2(g_probe, O_NONBLOCK | O_CLOEXEC) != 0) g_probe[0] = g_probe[1] = -1;
  g_mem_fd = open("/proc/self/mem", O_RDWR | O_CLOEXEC);
  // everything the fault handler needs is set up here, so that the handler itself never allocates
  if (!g_staging) g_staging = (char *)(g_copy.staging_alloc ? g_copy.staging_alloc(kStaging) : malloc(kStaging));
  install_handler();
  atexit(at_exit);
  g_installed = true;
}

Region *attach(void *host, size_t cap, void *dev) {
  char *h = (char *)host;
  // The first and last 64 bytes always stay in unprotected pages: free() writes its list links and footer there
  // when the host releases an array that was carved from the heap rather than mapped on its own.
  if (cap < 2 * kGuard) return nullptr;
  char *lo = (char *)(((uintptr_t)h + kGuard + g_page - 1) / g_page * g_page);
  char *hi = (char *)(((uintptr_t)h + cap - kGuard) / g_page * g_page);
  if (hi <= lo) return nullptr;
  // some mappings cannot change protection (driver-allocated page-locked memory, device files): leave those untracked
  if (mprotect(lo, g_page, PROT_NONE) != 0) return nullptr;
  if (mprotect(lo, g_page, PROT_READ | PROT_WRITE) != 0) fatal("vpic_b200: cannot restore page protection");
  Region *r = new Region;
  r->h = h; r->cap = cap; r->d = (char *)dev; r->lo = lo; r->hi = hi;
  r->base = (char *)((uintptr_t)lo / g_chunk * g_chunk);
  r->nchunks = ((size_t)(hi - r->base) + g_chunk - 1) / g_chunk;
  r->state = (uint8_t *)calloc(r->nchunks, 1);
  r->shadow = (char *)malloc((size_t)(lo - h) + (size_t)(h + cap - hi));
  Lock lk;
  int n = g_nregions.load();
  int slot = -1;
  for (int i = 0; i < n; i++) if (!g_regions[i]) { slot = i; break; }
  if (slot < 0) { if (n >= kMaxRegions) { free(r->state); delete r; return nullptr; } slot = n; }
  g_regions[slot] = r;
  if (slot == n) g_nregions.store(n + 1, std::memory_order_release);
  g_stats.regions++;
  return r;
}

void detach(Region *r, bool sync_host, uint64_t *d2h_bytes) {
  if (!r) return;
  {
    Lock lk;
    if (r->ndevice) {
      if (sync_host && still_protected(r)) fetch(r, 0, r->nchunks, d2h_bytes);
      else mprotect(r->lo, (size_t)(r->hi - r->lo), PROT_READ | PROT_WRITE);   // may fail if the host unmapped it: fine
    }
    const int n = g_nregions.load();
    for (int i = 0; i < n; i++) if (g_regions[i] == r) g_regions[i] = nullptr;
    g_stats.regions--;
  }
  for (int i = 0; i < g_nrefresh; i++) if (g_refresh[i] == r) g_refresh[i] = nullptr;
  free(r->shadow);
  free(r->state);
  delete r;
}

// shadow <- host for the ends the device just wrote (the host copy is complete when this runs, and the host program has
// not run since: after_sync() is called before the entry point returns)
static void finish_refresh(Region *r) {
  if (r->refresh_head) { memcpy(r->shadow, r->h, r->refresh_head_len); r->head_ok = r->refresh_head_len; }
  if (r->refresh_tail) { memcpy(r->shadow + (r->lo - r->h), r->hi, r->refresh_tail_len); r->tail_ok = r->refresh_tail_len; }
  r->refresh_head = r->refresh_tail = false;
  r->refresh_head_len = r->refresh_tail_len = 0;
}
// a refresh that never got its after_sync(): the host may have run since, so the shadow cannot be trusted
static void drop_refresh(Region *r) {
  if (r->refresh_head) r->head_ok = 0;
  if (r->refresh_tail) r->tail_ok = 0;
  r->refresh_head = r->refresh_tail = false;
  r->refresh_head_len = r->refresh_tail_len = 0;
}
static void note_refresh(Region *r, bool head, size_t len) {
  if (head) { r->refresh_head = true; if (len > r->refresh_head_len) r->refresh_head_len = len; r->head_ok = 0; }
  else { r->refresh_tail = true; if (len > r->refresh_tail_len) r->refresh_tail_len = len; r->tail_ok = 0; }
  if (!g_copy.d2h_async) { finish_refresh(r); return; }
  for (int i = 0; i < g_nrefresh; i++) if (g_refresh[i] == r) return;
  if (g_nrefresh < kMaxRegions) g_refresh[g_nrefresh++] = r; else drop_refresh(r);
}

void after_sync() {
  if (!g_nrefresh) return;
  Lock lk;
  for (int i = 0; i < g_nrefresh; i++) if (g_refresh[i]) finish_refresh(g_refresh[i]);
  g_nrefresh = 0;
}

void to_device(Region *r, size_t bytes, uint64_t *h2d_bytes) {
  if (bytes > r->cap) bytes = r->cap;
  if (!bytes) return;
  Lock lk;
  ensure_handler();
  if (r->ndevice && !still_protected(r)) {
    // remapped under us: every chunk is host-owned again, nothing to copy back
    mprotect(r->lo, (size_t)(r->hi - r->lo), PROT_READ | PROT_WRITE);
    memset(r->state, HOST, r->nchunks); r->ndevice = 0; g_stats.remaps++;
    r->head_ok = r->tail_ok = 0;
  }
  char *end = r->h + bytes;
  auto up = [&](char *a, char *b) {
    if (b <= a) return;
    if (g_copy.h2d(r->d + (a - r->h), a, (size_t)(b - a))) fatal("vpic_b200: host-to-device copy failed");
    if (h2d_bytes) *h2d_bytes += (uint64_t)(b - a);
  };
  // an unprotected end goes up only if the host changed it since host and device last agreed on it
  // (An end whose copy-back is still in flight — the array is touched twice inside one entry point — is uploaded
  // unconditionally: the copy engine reads the host bytes after that copy-back has landed, the CPU cannot compare yet.)
  const bool in_flight = r->refresh_head || r->refresh_tail;
  auto up_edge = [&](char *a, char *b, char *sh, size_t &ok) {
    if (b <= a) return;
    const size_t len = (size_t)(b - a);
    if (in_flight || !r->shadow) { up(a, b); ok = 0; return; }
    if (len <= ok && memcmp(a, sh, len) == 0) return;
    up(a, b);
    memcpy(sh, a, len); if (len > ok) ok = len;
  };
  drop_refresh(r);
  up_edge(r->h, end < r->lo ? end : r->lo, r->shadow, r->head_ok);     // head edge
  if (end > r->lo) {
    const size_t c1 = chunk_of(r, (end < r->hi ? end : r->hi) - 1) + 1;
    size_t c = 0;
    while (c < c1) {
      if (r->state[c] != HOST) { c++; continue; }
      size_t e = c;
      while (e < c1 && r->state[e] == HOST) e++;
      up(chunk_lo(r, c), chunk_hi(r, e - 1));                          // whole chunks: the tail past `bytes` rides along
      protect(r, c, e, PROT_NONE);
      for (size_t k = c; k < e; k++) r->state[k] = DEVICE;
      r->ndevice += e - c;
      c = e;
    }
  }
  if (end > r->hi) up_edge(r->hi, end, r->shadow ? r->shadow + (r->lo - r->h) : nullptr, r->tail_ok);   // tail edge
}

void device_wrote(Region *r, size_t bytes, uint64_t *d2h_bytes) {
  if (bytes > r->cap) bytes = r->cap;
  if (!bytes) return;
  Lock lk;
  char *end = r->h + bytes;
  auto down = [&](char *a, char *b) {
    if (b <= a) return;
    if ((g_copy.d2h_async ? g_copy.d2h_async : g_copy.d2h)(a, r->d + (a - r->h), (size_t)(b - a))) fatal("vpic_b200: device-to-host copy failed");
    if (d2h_bytes) *d2h_bytes += (uint64_t)(b - a);
  };
  // the copies below may be asynchronous (d2h_async): the shadow is re-read from the host in after_sync(), which the
  // caller runs once the device's work queue has drained, or right here when the copy is synchronous
  {
    char *b = end < r->lo ? end : r->lo;
    down(r->h, b);
    if (r->shadow && b > r->h) note_refresh(r, true, (size_t)(b - r->h));
  }
  if (end > r->lo) {
    // every chunk the device wrote must be device-owned (to_device ran first); one that a racing host thread took
    // back in between is handed to the device again — the host copy there is older than what was just written
    const size_t c1 = chunk_of(r, (end < r->hi ? end : r->hi) - 1) + 1;
    for (size_t c = 0; c < c1; c++)
      if (r->state[c] == HOST) { protect(r, c, c + 1, PROT_NONE); r->state[c] = DEVICE; r->ndevice++; }
  }
  if (end > r->hi) {
    down(r->hi, end);
    if (r->shadow) note_refresh(r, false, (size_t)(end - r->hi));
  }
}

void to_host(Region *r, size_t off, size_t bytes, uint64_t *d2h_bytes) {
  if (off >= r->cap) return;
  if (bytes > r->cap - off) bytes = r->cap - off;
  char *a = r->h + off, *b = a + bytes;
  if (a < r->lo) a = r->lo;
  if (b > r->hi) b = r->hi;
  if (b <= a) return;
  Lock lk;
  if (!r->ndevice) return;
  fetch(r, chunk_of(r, a), chunk_of(r, b - 1) + 1, d2h_bytes);
}

void forget_device(Region *r) {
  Lock lk;
  r->head_ok = r->tail_ok = 0; r->refresh_head = r->refresh_tail = false;
  if (!r->ndevice) return;
  mprotect(r->lo, (size_t)(r->hi - r->lo), PROT_READ | PROT_WRITE);
  memset(r->state, HOST, r->nchunks); r->ndevice = 0;
}

int host_access(const void *p, size_t n) {
  if (!n || !g_nregions.load(std::memory_order_acquire)) return 0;
  const char *a = (const char *)p, *b = a + n;
  int touched = 0;
  {
    // Without the lock first: this runs under every read(2)/write(2) of the process (dropin.cu interposes them), also
    // on threads of the CUDA runtime while another thread holds the lock and waits for the device.  A buffer that
    // overlaps no tracked array must never wait here.  (Regions are only added or removed between entry points.)
    bool any = false;
    const int nr0 = g_nregions.load(std::memory_order_acquire);
    for (int i = 0; i < nr0 && !any; i++) { const Region *r = g_regions[i]; any = r && b > r->lo && a < r->hi; }
    if (!any) return 0;
  }
  Lock lk;
  const int nr = g_nregions.load();
  for (int i = 0; i < nr; i++) {
    Region *r = g_regions[i];
    if (!r || !r->ndevice || b <= r->lo || a >= r->hi) continue;
    const char *x = a < r->lo ? r->lo : a, *y = b > r->hi ? r->hi : b;
    uint64_t bytes = 0;
    fetch(r, chunk_of(r, x), chunk_of(r, y - 1) + 1, &bytes);
    g_stats.fault_bytes += bytes;
    touched++;
  }
  return touched;
}

bool device_owns(Region *r, size_t off) {
  if (!r || off >= r->cap) return false;
  char *a = r->h + off;
  if (a < r->lo || a >= r->hi) return false;            // the edges are always host-owned
  Lock lk;
  return r->ndevice && r->state[chunk_of(r, a)] != HOST;
}

bool all_device(Region *r, size_t bytes) {
  if (!r) return false;
  if (bytes > r->cap) bytes = r->cap;
  char *end = r->h + bytes;
  if (end <= r->lo) return true;
  Lock lk;
  if (r->ndevice && !still_protected(r)) return false;
  const size_t c1 = chunk_of(r, (end < r->hi ? end : r->hi) - 1) + 1;
  for (size_t c = 0; c < c1; c++) if (r->state[c] != DEVICE) return false;
  return true;
}

void set_device(Region *r, void *dev) {
  if (!r) return;
  Lock lk;
  r->d = (char *)dev;
}

void edges(Region *r, size_t *head_end, size_t *tail_begin) {
  *head_end = (size_t)(r->lo - r->h);
  *tail_begin = (size_t)(r->hi - r->h);
}

int active() { return (int)g_stats.regions; }

Stats stats() { return g_stats; }

}  // namespace vpb_lazy
