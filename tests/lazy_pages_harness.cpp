// CPU-only self-test of vpic_b200/csrc/lazy_pages.cpp (the page-protection tracker behind VPB_MODE_AUTO).
// The "device" is a second host buffer and the copies are memcpy, so every state transition, the fault handler,
// the edge handling, the remap detection and concurrent faulting threads can be exercised without a GPU.
// Built by `make -C vpic_b200/csrc lazy_test`, run by tests/test_host.py.  Exit code 0 = pass; otherwise the
// failing line is printed.
#include "lazy_pages.h"
#include <vector>

#include <pthread.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>

static uint64_t g_h2d_calls = 0, g_d2h_calls = 0;
struct Pending { void *h; const void *d; size_t n; };
static std::vector<Pending> g_pending_copies;
// one in-order work queue, like a CUDA stream: a copy that is issued later executes after the queued copy-backs
static void drain_queue() { for (const Pending &p : g_pending_copies) memcpy(p.h, p.d, p.n); g_pending_copies.clear(); }
static int fake_h2d(void *d, const void *h, size_t n) { drain_queue(); memcpy(d, h, n); g_h2d_calls++; return 0; }
static int fake_d2h(void *h, const void *d, size_t n) { drain_queue(); memcpy(h, d, n); g_d2h_calls++; return 0; }
// asynchronous copy-back of the array ends (Copier::d2h_async): the bytes only arrive when the "work queue" drains, which
// an entry point waits for once before it returns (finish_entry in dropin.cu) and then tells the tracker (after_sync)
static int fake_d2h_async(void *h, const void *d, size_t n) { g_pending_copies.push_back({h, d, n}); g_d2h_calls++; return 0; }
static void entry_end() { drain_queue(); vpb_lazy::after_sync(); }
static void fake_fatal(const char *m) { fprintf(stderr, "FATAL: %s\n", m); _exit(3); }

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "lazy_pages_test: line %d: %s\n", __LINE__, #c); return __LINE__; } } while (0)

static const size_t kPage = 4096, kChunk = 4 * 4096;

struct Arr { char *map, *h, *d; size_t cap; };
// host array at an odd offset inside its mapping (like MALLOC_ALIGNED's 128-byte alignment), ending mid-page
static Arr make(size_t pages, size_t head, size_t tail_cut) {
  Arr a;
  a.map = (char *)mmap(nullptr, pages * kPage, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  a.h = a.map + head; a.cap = pages * kPage - head - tail_cut;
  a.d = (char *)malloc(a.cap);
  memset(a.d, 0xEE, a.cap);
  return a;
}
static void fill(char *p, size_t n, int seed) { for (size_t i = 0; i < n; i++) p[i] = (char)((i * 31 + seed) & 0x7f); }
static bool same(const char *p, size_t n, int seed, int add) {
  for (size_t i = 0; i < n; i++) if (p[i] != (char)((((i * 31 + seed) & 0x7f) + add) & 0xff)) return false;
  return true;
}
static void device_kernel(Arr &a, size_t n) { for (size_t i = 0; i < n; i++) a.d[i] = (char)(a.d[i] + 1); }   // "+1 everywhere"

struct Reader { const char *p; size_t n; long sum; };
static void *reader_main(void *arg) { Reader *r = (Reader *)arg; long s = 0; for (size_t i = 0; i < r->n; i++) s += r->p[i]; r->sum = s; return nullptr; }

static int run() {
  vpb_lazy::Copier cp = {fake_h2d, fake_d2h, fake_fatal, nullptr, nullptr, nullptr};
  vpb_lazy::init(cp, kChunk);
  uint64_t h2d = 0, d2h = 0;

  // --- 1. upload, device write, host read faults the data back -------------------------------------------------
  Arr a = make(40, 128, 1000);
  fill(a.h, a.cap, 1);
  vpb_lazy::Region *r = vpb_lazy::attach(a.h, a.cap, a.d);
  CHECK(r != nullptr && vpb_lazy::active() == 1);
  vpb_lazy::to_device(r, a.cap, &h2d);
  CHECK(h2d == a.cap);
  CHECK(memcmp(a.d, a.h, 128) == 0);                         // (only the unprotected head can be compared directly)
  device_kernel(a, a.cap);
  vpb_lazy::device_wrote(r, a.cap, &d2h);
  CHECK(d2h == (kPage - 128) + (kPage - 1000));              // the two edges came back eagerly
  vpb_lazy::Stats s0 = vpb_lazy::stats();
  CHECK(same(a.h, a.cap, 1, 1));                             // reads every byte: faults chunk by chunk
  vpb_lazy::Stats s1 = vpb_lazy::stats();
  CHECK(s1.faults > s0.faults && s1.faults - s0.faults <= 10);
  CHECK(s1.fault_bytes - s0.fault_bytes == 38 * kPage);      // exactly the whole pages, once
  CHECK(same(a.h, a.cap, 1, 1));
  CHECK(vpb_lazy::stats().faults == s1.faults);              // host-owned now: no more faults

  // --- 2. steady state: nothing host-touched means nothing copied ---------------------------------------------
  h2d = d2h = 0;
  vpb_lazy::to_device(r, a.cap, &h2d);                       // every whole page is host-owned after the read: they go up
  CHECK(h2d == 38 * kPage);                                  // ... the two ends do not: the host only read them
  for (int step = 0; step < 5; step++) {
    h2d = d2h = 0;
    vpb_lazy::to_device(r, a.cap, &h2d);
    device_kernel(a, a.cap);
    vpb_lazy::device_wrote(r, a.cap, &d2h);
    CHECK(h2d == 0 && d2h == (kPage - 128) + (kPage - 1000));       // the ends come back, nothing goes up
  }
  // --- 2b. the unprotected ends: an end goes up again exactly when the host changed it ----------------------------
  {
    Arr e = make(12, 128, 1000);
    fill(e.h, e.cap, 9);
    vpb_lazy::Region *re = vpb_lazy::attach(e.h, e.cap, e.d);
    const size_t head = kPage - 128, tail = kPage - 1000;
    h2d = d2h = 0;
    vpb_lazy::to_device(re, e.cap, &h2d);
    CHECK(h2d == e.cap);
    device_kernel(e, e.cap);
    vpb_lazy::device_wrote(re, e.cap, &d2h);
    CHECK(d2h == head + tail);
    e.h[5] = (char)(e.h[5] + 3);                             // host edit at the head (no fault: the page is not protected)
    h2d = d2h = 0;
    vpb_lazy::to_device(re, e.cap, &h2d);
    CHECK(h2d == head && e.d[5] == e.h[5]);                  // the head went up, the tail did not
    device_kernel(e, e.cap);
    vpb_lazy::device_wrote(re, e.cap, &d2h);
    CHECK(e.h[5] == e.d[5] && d2h == head + tail);
    e.h[e.cap - 3] = 77;                                     // ... and at the tail
    h2d = 0;
    vpb_lazy::to_device(re, e.cap, &h2d);
    CHECK(h2d == tail && e.d[e.cap - 3] == 77);
    h2d = 0;
    vpb_lazy::to_device(re, e.cap, &h2d);                    // a second entry point without a device write in between
    CHECK(h2d == 0);
    // a shorter extent compares (and uploads) only what it covers; the rest of the tail stays as the device has it
    e.h[e.cap - 3] = 78;
    h2d = 0;
    vpb_lazy::to_device(re, e.cap - 500, &h2d);
    CHECK(h2d == 0 && e.d[e.cap - 3] == 77);
    vpb_lazy::to_device(re, e.cap, &h2d);
    CHECK(h2d == tail && e.d[e.cap - 3] == 78);
    // the device writes the ends, the host edits them afterwards: the edit wins at the next upload
    device_kernel(e, e.cap);
    vpb_lazy::device_wrote(re, e.cap, &d2h);
    const char dev_val = e.h[7];
    e.h[7] = (char)(dev_val + 5);
    h2d = 0;
    vpb_lazy::to_device(re, e.cap, &h2d);
    CHECK(h2d == head && e.d[7] == (char)(dev_val + 5));
    // forget_device: the host copy is declared current, so everything goes up again, the ends included
    vpb_lazy::forget_device(re);
    h2d = 0;
    vpb_lazy::to_device(re, e.cap, &h2d);
    CHECK(h2d == e.cap);
    vpb_lazy::detach(re, true, &d2h);
    munmap(e.map, 12 * kPage); free(e.d);
  }
  // --- 3. a sparse host write (inject_particle) moves one chunk, and the write survives the next device pass ---
  s0 = vpb_lazy::stats();
  a.h[20 * kPage + 7] = 99;
  s1 = vpb_lazy::stats();
  CHECK(s1.faults == s0.faults + 1 && s1.fault_bytes - s0.fault_bytes == kChunk);
  CHECK(a.h[20 * kPage + 8] == (char)(((((20 * kPage + 8) * 31 + 1) & 0x7f) + 6) & 0xff));   // neighbours are current
  h2d = 0;
  vpb_lazy::to_device(r, a.cap, &h2d);
  CHECK(h2d == kChunk);
  device_kernel(a, a.cap);
  vpb_lazy::device_wrote(r, a.cap, &d2h);
  CHECK(a.h[20 * kPage + 7] == 100);

  // --- 4. partial extents: the device only uses the first part; the host's tail is preserved -------------------
  {
    Arr b = make(24, 256, 0);
    fill(b.h, b.cap, 5);
    vpb_lazy::Region *rb = vpb_lazy::attach(b.h, b.cap, b.d);
    const size_t live = 9 * kPage + 100;
    h2d = 0;
    vpb_lazy::to_device(rb, live, &h2d);
    device_kernel(b, live);
    vpb_lazy::device_wrote(rb, live, &d2h);
    CHECK(same(b.h, live, 5, 1));
    for (size_t i = live; i < b.cap; i++) CHECK(b.h[i] == (char)((i * 31 + 5) & 0x7f));       // untouched tail
    // to_host on a sub-range, forget_device, detach with sync
    vpb_lazy::to_device(rb, b.cap, &h2d);
    device_kernel(b, b.cap);
    vpb_lazy::device_wrote(rb, b.cap, &d2h);
    d2h = 0;
    vpb_lazy::to_host(rb, 5 * kPage, 2 * kPage, &d2h);
    CHECK(d2h >= 2 * kPage && d2h <= 2 * kChunk);
    vpb_lazy::detach(rb, true, &d2h);
    CHECK(vpb_lazy::active() == 1);
    for (size_t i = 0; i < live; i++) CHECK(b.h[i] == (char)((((i * 31 + 5) & 0x7f) + 2) & 0xff));
    for (size_t i = live; i < b.cap; i++) CHECK(b.h[i] == (char)((((i * 31 + 5) & 0x7f) + 1) & 0xff));
    munmap(b.map, 24 * kPage); free(b.d);
  }

  // --- 5. host_access (what the fwrite/fread interposers call): no fault, range becomes readable by the kernel --
  vpb_lazy::to_device(r, a.cap, &h2d);
  {
    int fd[2]; CHECK(pipe(fd) == 0);
    ssize_t w = write(fd[1], a.h + 8 * kPage, 64);
    CHECK(w < 0);                                            // EFAULT: device-owned pages are invisible to syscalls
    s0 = vpb_lazy::stats();
    CHECK(vpb_lazy::host_access(a.h + 8 * kPage, 64) == 1);
    w = write(fd[1], a.h + 8 * kPage, 64);
    CHECK(w == 64 && vpb_lazy::stats().faults == s0.faults);
    close(fd[0]); close(fd[1]);
  }

  // --- 6. several host threads fault on the same array at once (the reference's pipelines) --------------------
  vpb_lazy::to_device(r, a.cap, &h2d);
  device_kernel(a, a.cap);
  vpb_lazy::device_wrote(r, a.cap, &d2h);
  {
    long expect = 0;
    std::vector<char> copy(a.d, a.d + a.cap);
    for (size_t i = 0; i < a.cap; i++) expect += copy[i];
    pthread_t th[4]; Reader rd[4];
    for (int t = 0; t < 4; t++) { rd[t] = {a.h, a.cap, 0}; pthread_create(&th[t], nullptr, reader_main, &rd[t]); }
    for (int t = 0; t < 4; t++) { pthread_join(th[t], nullptr); CHECK(rd[t].sum == expect); }
  }

  // --- 7. the host frees the array and the allocator maps fresh memory at the same address ---------------------
  vpb_lazy::to_device(r, a.cap, &h2d);
  munmap(a.map, 40 * kPage);
  char *again = (char *)mmap(a.map, 40 * kPage, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_FIXED, -1, 0);
  CHECK(again == a.map);
  fill(a.h, a.cap, 9);                                       // no faults: fresh pages
  h2d = 0;
  vpb_lazy::to_device(r, a.cap, &h2d);
  CHECK(vpb_lazy::stats().remaps == 1 && h2d == a.cap);      // noticed, everything uploaded again
  CHECK(memcmp(a.d, a.h, 128) == 0);
  device_kernel(a, a.cap);
  vpb_lazy::device_wrote(r, a.cap, &d2h);
  CHECK(same(a.h, a.cap, 9, 1));

  // --- 8. forget_device (vpic_b200_invalidate): host copy wins without a copy back -----------------------------
  vpb_lazy::to_device(r, a.cap, &h2d);
  device_kernel(a, a.cap);
  vpb_lazy::device_wrote(r, a.cap, &d2h);
  vpb_lazy::forget_device(r);
  s0 = vpb_lazy::stats();
  volatile char c = a.h[10 * kPage]; (void)c;
  CHECK(vpb_lazy::stats().faults == s0.faults);
  vpb_lazy::detach(r, false, nullptr);
  CHECK(vpb_lazy::active() == 0);

  // --- 8b. somebody installs a SIGSEGV handler after us: the next to_device takes the signal back and chains ------
  {
    Arr t = make(12, 128, 0);
    fill(t.h, t.cap, 3);
    vpb_lazy::Region *rt = vpb_lazy::attach(t.h, t.cap, t.d);
    signal(SIGSEGV, SIG_DFL);                                  // e.g. a runtime resetting handlers
    vpb_lazy::to_device(rt, t.cap, &h2d);
    device_kernel(t, t.cap);
    vpb_lazy::device_wrote(rt, t.cap, &d2h);
    CHECK(same(t.h, t.cap, 3, 1));                             // would crash here without the re-install
    vpb_lazy::detach(rt, false, nullptr);
  }

  // --- 9. an array with no whole page inside is not tracked -----------------------------------------------------
  {
    Arr t = make(2, 100, kPage + 100);
    CHECK(vpb_lazy::attach(t.h, t.cap, t.d) == nullptr);
  }
  printf("lazy_pages_test: ok (%llu faults, %llu h2d calls, %llu d2h calls)\n",
         (unsigned long long)vpb_lazy::stats().faults, (unsigned long long)g_h2d_calls, (unsigned long long)g_d2h_calls);
  return 0;
}

// Randomised model check: a long random sequence of entry-point calls with random extents, host reads, host writes
// and kernel-side accesses, each compared with a byte-exact model of what the program must observe.
static uint64_t rng_state = 1;
static uint32_t rnd() { rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(rng_state >> 33); }

static int stress(int iterations, uint64_t seed, bool async_ends) {
  rng_state = seed * 2654435761u + 17;
  vpb_lazy::Copier cp = {fake_h2d, fake_d2h, fake_fatal, nullptr, nullptr, async_ends ? fake_d2h_async : nullptr};
  vpb_lazy::init(cp, kChunk);
  const size_t pages = 64 + rnd() % 64;
  Arr a = make(pages, 128 * (1 + rnd() % 20), rnd() % 3000);
  std::vector<char> truth(a.cap);
  for (size_t i = 0; i < a.cap; i++) { truth[i] = (char)(rnd() & 0x7f); a.h[i] = truth[i]; }
  vpb_lazy::Region *r = vpb_lazy::attach(a.h, a.cap, a.d);
  CHECK(r != nullptr);
  uint64_t h2d = 0, d2h = 0;
  for (int it = 0; it < iterations; it++) {
    const uint32_t op = rnd() % 100;
    size_t off = rnd() % a.cap, len = 1 + rnd() % (op % 7 == 0 ? a.cap : 3 * kPage);
    if (off + len > a.cap) len = a.cap - off;
    if (op < 30) {                                            // an entry point: device reads and rewrites [0, ext)
      const size_t ext = (op % 3 == 0) ? a.cap : 1 + rnd() % a.cap;
      vpb_lazy::to_device(r, ext, &h2d);
      CHECK(memcmp(a.d, truth.data(), ext) == 0);            // the device sees exactly what the program last had
      for (size_t i = 0; i < ext; i++) { a.d[i] = (char)(a.d[i] + 1); truth[i] = (char)(truth[i] + 1); }
      vpb_lazy::device_wrote(r, ext, &d2h);
      if (async_ends && op % 2) {                             // an entry point that touches the array twice before it returns
        const size_t ext2 = 1 + rnd() % a.cap;                // (no wait in between: the copy-backs of the ends are in flight)
        vpb_lazy::to_device(r, ext2, &h2d);
        drain_queue();
        CHECK(memcmp(a.d, truth.data(), ext2) == 0);
        for (size_t i = 0; i < ext2; i++) { a.d[i] = (char)(a.d[i] + 1); truth[i] = (char)(truth[i] + 1); }
        vpb_lazy::device_wrote(r, ext2, &d2h);
      }
      entry_end();
    } else if (op < 40) {                                     // an entry point that only reads (e.g. interpolators in advance_p)
      const size_t ext = 1 + rnd() % a.cap;
      vpb_lazy::to_device(r, ext, &h2d);
      CHECK(memcmp(a.d, truth.data(), ext) == 0);
    } else if (op < 65) {                                     // host reads a range
      CHECK(memcmp(a.h + off, truth.data() + off, len) == 0);
    } else if (op < 85) {                                     // host writes a range
      for (size_t i = 0; i < len; i++) { const char v = (char)(rnd() & 0x7f); a.h[off + i] = v; truth[off + i] = v; }
    } else if (op < 92) {                                     // kernel-side read (fwrite): host_access first, then a syscall
      vpb_lazy::host_access(a.h + off, len);
      int fd[2]; CHECK(pipe(fd) == 0);
      const size_t n = len < 4096 ? len : 4096;
      std::vector<char> got(n);
      CHECK(write(fd[1], a.h + off, n) == (ssize_t)n && read(fd[0], got.data(), n) == (ssize_t)n);
      CHECK(memcmp(got.data(), truth.data() + off, n) == 0);
      close(fd[0]); close(fd[1]);
    } else if (op < 96) {                                     // explicit sync of a range (vpic_b200_sync_to_host)
      vpb_lazy::to_host(r, off, len, &d2h);
      CHECK(memcmp(a.h + off, truth.data() + off, len) == 0);
    } else {                                                  // two host threads read the whole array at once
      long expect = 0;
      for (size_t i = 0; i < a.cap; i++) expect += truth[i];
      pthread_t th[2]; Reader rd[2];
      for (int t = 0; t < 2; t++) { rd[t] = {a.h, a.cap, 0}; pthread_create(&th[t], nullptr, reader_main, &rd[t]); }
      for (int t = 0; t < 2; t++) { pthread_join(th[t], nullptr); CHECK(rd[t].sum == expect); }
    }
  }
  vpb_lazy::detach(r, true, &d2h);
  CHECK(memcmp(a.h, truth.data(), a.cap) == 0);               // after detach the host holds everything
  printf("lazy_pages_test: stress ok (%d operations%s, seed %llu, %llu faults, %.1f MB up, %.1f MB down)\n", iterations,
         async_ends ? ", asynchronous array ends" : "",
         (unsigned long long)seed, (unsigned long long)vpb_lazy::stats().faults, h2d / 1e6, (d2h + vpb_lazy::stats().fault_bytes) / 1e6);
  return 0;
}

// Throughput of the staged fetch path (device -> staging -> /proc/self/mem -> protected pages), the route pageable
// host arrays take; the copies are memcpy here, so this bounds the host-side overhead of that route.
static int bandwidth(size_t mb) {
  vpb_lazy::Copier cp = {fake_h2d, fake_d2h, fake_fatal, nullptr, nullptr, nullptr};
  vpb_lazy::init(cp, 2u << 20);
  const size_t pages = mb * 256;
  Arr a = make(pages, 128, 0);
  memset(a.h, 1, a.cap);
  vpb_lazy::Region *r = vpb_lazy::attach(a.h, a.cap, a.d);
  uint64_t h2d = 0, d2h = 0;
  vpb_lazy::to_device(r, a.cap, &h2d);
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  long s = 0;
  for (size_t i = 0; i < a.cap; i += 4096) s += a.h[i];       // one touch per page: faults drive the fetches
  clock_gettime(CLOCK_MONOTONIC, &t1);
  const double sec = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
  printf("lazy_pages_test: fetched %.0f MB in %llu faults, %.2f GB/s through the staged path (checksum %ld)\n",
         vpb_lazy::stats().fault_bytes / 1e6, (unsigned long long)vpb_lazy::stats().faults,
         vpb_lazy::stats().fault_bytes / sec / 1e9, s);
  vpb_lazy::detach(r, false, &d2h);
  return 0;
}

int main(int argc, char **argv) {
  if (argc > 2 && !strcmp(argv[1], "--bandwidth")) return bandwidth((size_t)atoi(argv[2]));
  if (argc > 3 && !strcmp(argv[1], "--stress")) return stress(atoi(argv[2]), strtoull(argv[3], nullptr, 10), false) ? 1 : 0;
  if (argc > 3 && !strcmp(argv[1], "--stress-async")) return stress(atoi(argv[2]), strtoull(argv[3], nullptr, 10), true) ? 1 : 0;
  const int rc = run();
  if (rc) return 1;
  if (argc > 1 && !strcmp(argv[1], "--crash")) {             // a genuine wild access must still kill the process
    volatile int *bad = (volatile int *)16; *bad = 1;
  }
  return 0;
}
