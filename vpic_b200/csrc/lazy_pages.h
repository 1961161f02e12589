// Page-protected lazy coherence between a host array and its device mirror (VPB_MODE_AUTO of the drop-in layer).
//
// The reference's host program owns every array (MALLOC_ALIGNED, src/util/util_base.h:150-170) and reads or writes
// them whenever it likes between hot-path calls: decks poke sp->p[n] and field(x,y,z), boundary_p walks sp->pm,
// dumps and checkpoints fwrite whole arrays (src/vpic/dump.cc:217, src/util/io/StandardIOPolicy.h:133-145).  Copying
// every array across PCIe on every call (VPB_MODE_COHERENT) keeps that contract but costs ~80 bytes of PCIe traffic
// per particle per step.  Here the contract is kept by the MMU instead:
//
//   * the whole pages inside [host, host+cap) are cut into chunks (2 MB, on absolute 2 MB boundaries so transparent
//     huge pages survive); every chunk is owned by the HOST (pages read/write, device copy stale) or by the DEVICE
//     (pages PROT_NONE, host copy stale);
//   * a drop-in entry point that needs the array uploads the host-owned chunks and turns them device-owned;
//   * the first host access to a device-owned chunk faults; the SIGSEGV handler copies the chunk back, unprotects
//     it, and the faulting instruction is retried.  Sequential scans grow the window that is fetched per fault;
//   * the partial pages at both ends of the array are shared with foreign heap data and are never protected: those
//     bytes ("edges", < 8 KB per array) come back to the host after every call that wrote the array, and go up again
//     whenever the host changed them — the tracker keeps a copy of what host and device last agreed on and compares;
//   * kernel-side accesses do not fault (write(2) on PROT_NONE memory fails with EFAULT), so the library interposes
//     fwrite/fread, the only I/O calls the reference makes on these arrays, and exports vpic_b200_host_access()
//     for anything else (MPI on device-owned memory).
//
// The handler calls mprotect (async-signal-safe) and the d2h callback (a CUDA copy, which is not).  That is sound here
// because the signal is synchronous: it is raised by the faulting load or store of host code that, by construction,
// is never inside the CUDA runtime, malloc or this tracker (none of them touches the protected interior pages, and
// the tracker's own critical sections are recognised by thread id and passed on to the previous handler), so no lock
// the callback needs can be held by the interrupted code.  Faults from several host threads are serialised by a spin
// lock; a fault on an address no tracked array owns goes to the handler that was installed before (or the default).
//
// Host-only code with the copies behind function pointers, so the state machine is testable without a GPU
// (tests/lazy_pages_harness.cpp).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace vpb_lazy {

struct Copier {
  int (*h2d)(void *dev, const void *host, size_t n);   // may be asynchronous on the device's work queue
  int (*d2h)(void *host, const void *dev, size_t n);   // complete on return, ordered after all earlier device work
  void (*fatal)(const char *msg);                      // does not return
  // Optional: copy into host pages that are still PROT_NONE (a DMA engine writing page-locked memory does not look
  // at CPU page protections).  Returns nonzero when it cannot (pageable memory); the tracker then stages the data
  // and writes it through /proc/self/mem, which also bypasses the protection.  Either way the pages only become
  // accessible after they hold the data, so a second host thread can never read a half-filled chunk.
  int (*d2h_protected)(void *host, const void *dev, size_t n);
  void *(*staging_alloc)(size_t n);                    // optional (page-locked) staging buffer; malloc when null
  // Optional: like d2h but may return before the bytes have arrived; the caller of device_wrote() waits for the device's
  // work queue once, after all the arrays of an entry point (used for the unprotected array ends only).
  int (*d2h_async)(void *host, const void *dev, size_t n);
};

struct Region;

struct Stats { uint64_t faults, fault_bytes, remaps, regions; };

void init(const Copier &c, size_t chunk_bytes);        // installs the SIGSEGV handler (once)
Region *attach(void *host, size_t cap, void *dev);     // nullptr when no whole page lies inside [host, host+cap)
void detach(Region *r, bool sync_host, uint64_t *d2h_bytes);
void to_device(Region *r, size_t bytes, uint64_t *h2d_bytes);     // device copy of [0,bytes) made current
void device_wrote(Region *r, size_t bytes, uint64_t *d2h_bytes);  // edge bytes inside [0,bytes) copied back
void to_host(Region *r, size_t off, size_t bytes, uint64_t *d2h_bytes);
void forget_device(Region *r);                         // the host copy is declared current everywhere
void after_sync();                                     // the copy-backs queued by device_wrote() (d2h_async) have completed
int host_access(const void *p, size_t n);              // [p,p+n) made host-owned in every attached region
bool device_owns(Region *r, size_t off);               // is the byte at offset `off` currently owned by the device?
bool all_device(Region *r, size_t bytes);               // every whole page inside [0,bytes) is device-owned (the host has not touched it)
void set_device(Region *r, void *dev);                 // the device copy now lives at `dev` (same contents, same capacity)
// the unprotected ends of the array: bytes [0, *head_end) and [*tail_begin, cap) are never device-owned
void edges(Region *r, size_t *head_end, size_t *tail_begin);
int active();                                          // number of attached regions
Stats stats();

}  // namespace vpb_lazy
