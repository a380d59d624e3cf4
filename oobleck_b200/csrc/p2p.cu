// Inter-stage activation / gradient exchange as direct NVLink peer-memory writes (replaces the blocking
// ncclSend/ncclRecv per tuple member of oobleck/execution/pipeline.py:270-286, 331-333, 383-387, 391-427).
//
// Every rank owns one cudaMalloc'ed "mailbox" per neighbour, mapped into the neighbour with CUDA IPC:
//
//     [ flags[64] | acks[64] | counters[8] | abort | pad ][ ring: nslots x slot_bytes ]
//       ^ written by the peer            ^ local scratch            ^ written by the peer (payload)
//
//   send(n):  (sender's stream) wait local acks[slot] >= n - nslots   -- the peer has drained the slot
//             copy payload -> peer ring[slot] with 128-bit stores over NVLink, __threadfence_system
//             last CTA: st.release.sys peer flags[slot] = n
//   recv(n):  (receiver's stream) spin ld.acquire.sys local flags[slot] >= n, copy ring[slot] -> destination,
//             last CTA: st.release.sys peer acks[slot] = n
//
// No NCCL rendezvous per message, no host synchronisation: both kernels are ordinary stream work, so the copy
// overlaps the 1F1B compute on the dedicated copy streams the host side gives them.
#include "../../include/oobleck_b200.h"
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "kernels.h"

using namespace oob;

namespace {

constexpr int P2P_MAX_SLOTS = 64;
struct MailboxHeader {
  unsigned flags[P2P_MAX_SLOTS];
  unsigned acks[P2P_MAX_SLOTS];
  unsigned counters[8];
  unsigned abort;
  unsigned pad[119];
};
static_assert(sizeof(MailboxHeader) == 1024, "mailbox header must stay 1 KiB");

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Watchdog: a spin that outlives this many ns (peer died before anyone called oob_p2p_abort, or a rendezvous bug)
// gives up with status 2, so a lost neighbour can never wedge the GPU.  OOB_P2P_TIMEOUT_S.
__device__ unsigned long long g_spin_timeout_ns = 120ull * 1000000000ull;

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Host-visible control block of a mailbox (pinned, mapped): `abort` is written by the HOST with a plain store --
// no stream, no CUDA call, so the listener thread can release a GPU whose copy streams are all blocked behind
// spinning kernels -- and polled by the waiting kernel; `status` is written by the device when a wait gave up
// (1 = aborted, 2 = watchdog) and read by the host after the step (oob_p2p_status).
struct HostCtrl {
  volatile unsigned abort;
  volatile unsigned status;
};

// One warp waits until *word >= target (wrap-safe).  On abort / timeout it latches the DEVICE abort word of the
// mailbox: every later copy kernel of this link then skips its copy and its publish, so a lost or late peer can never
// turn into silently corrupted activations or a corrupted flag / ack sequence.
__global__ void p2p_wait_kernel(const unsigned* word, unsigned target, unsigned* abort_dev, HostCtrl* ctrl) {
  if (threadIdx.x != 0) return;
  unsigned ns = 32;
  const unsigned long long t0 = globaltimer_ns();
  while ((int)(ld_acquire_sys(word) - target) < 0) {
    unsigned why = 0;
    if (ld_acquire_sys(abort_dev)) why = 1;
    else if (ns >= 1024) {   // slow path only: the host word costs a PCIe round trip per poll
      if (ctrl->abort) why = 1;
      else if (globaltimer_ns() - t0 > g_spin_timeout_ns) why = 2;
    }
    if (why) {
      atomicExch(abort_dev, why);
      ctrl->status = why;
      __threadfence_system();
      return;
    }
    __nanosleep(ns);
    if (ns < 1024) ns <<= 1;
  }
}

__device__ __forceinline__ void copy_bytes(const char* src, char* dst, long bytes) {
  const long n16 = bytes >> 4;
  const int4* s = reinterpret_cast<const int4*>(src);
  int4* d = reinterpret_cast<int4*>(dst);
  const long stride = (long)gridDim.x * blockDim.x;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  // 4 independent 128-bit accesses in flight per thread
  for (; i + 3 * stride < n16; i += 4 * stride) {
    int4 a = s[i], b = s[i + stride], c = s[i + 2 * stride], e = s[i + 3 * stride];
    d[i] = a; d[i + stride] = b; d[i + 2 * stride] = c; d[i + 3 * stride] = e;
  }
  for (; i < n16; i += stride) d[i] = s[i];
  if (blockIdx.x == 0 && threadIdx.x < (bytes & 15)) dst[(n16 << 4) + threadIdx.x] = src[(n16 << 4) + threadIdx.x];
}

// signal: after all CTAs finished their part, the last one publishes `value` at `word` (system scope)
__device__ __forceinline__ void publish_when_all_done(unsigned* counter, unsigned* word, unsigned value) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(counter, 1u);
    if (done == gridDim.x - 1) {
      *counter = 0;
      __threadfence_system();
      if (word) st_release_sys(word, value);
    }
  }
}

// The copy itself: ordered behind p2p_wait_kernel by the stream (one resident spinning WARP per pending message instead
// of round 1's up-to-64 spinning CTAs next to the persistent GEMM).  An aborted link copies and publishes nothing.
__global__ void __launch_bounds__(256)
p2p_copy_kernel(const char* __restrict__ src, char* __restrict__ dst, long bytes, unsigned* publish_word, unsigned seq,
                unsigned* counter, const unsigned* abort_dev) {
  if (ld_acquire_sys(abort_dev)) return;
  copy_bytes(src, dst, bytes);
  publish_when_all_done(counter, publish_word, seq);
}

struct CtrlEntry { HostCtrl* host; HostCtrl* dev; };
std::mutex g_ctrl_mu;
std::unordered_map<const void*, CtrlEntry> g_ctrl;
bool ctrl_of(const void* mailbox, CtrlEntry* out) {
  std::lock_guard<std::mutex> lk(g_ctrl_mu);
  auto it = g_ctrl.find(mailbox);
  if (it == g_ctrl.end()) return false;
  *out = it->second;
  return true;
}

}  // namespace

extern "C" {

long oob_p2p_header_bytes(void) { return (long)sizeof(MailboxHeader); }

int oob_p2p_alloc(long ring_bytes, void** mailbox, void* ipc_handle_out) {
  OOB_CHECK(mailbox && ipc_handle_out && ring_bytes >= 0, "oob_p2p_alloc: bad arguments");
  void* p = nullptr;
  const size_t total = sizeof(MailboxHeader) + (size_t)ring_bytes;
  OOB_CUDA_OK(cudaMalloc(&p, total));
  OOB_CUDA_OK(cudaMemset(p, 0, sizeof(MailboxHeader)));
  cudaIpcMemHandle_t h;
  OOB_CUDA_OK(cudaIpcGetMemHandle(&h, p));
  memcpy(ipc_handle_out, &h, sizeof(h));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  CtrlEntry ce{};
  OOB_CUDA_OK(cudaHostAlloc(reinterpret_cast<void**>(&ce.host), 64, cudaHostAllocMapped | cudaHostAllocPortable));
  ce.host->abort = 0;
  ce.host->status = 0;
  OOB_CUDA_OK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&ce.dev), ce.host, 0));
  {
    std::lock_guard<std::mutex> lk(g_ctrl_mu);
    g_ctrl[p] = ce;
  }
  *mailbox = p;
  return 0;
}

int oob_p2p_open(const void* ipc_handle, void** peer_mailbox) {
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof(h));
  OOB_CUDA_OK(cudaIpcOpenMemHandle(peer_mailbox, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int oob_p2p_close(void* peer_mailbox) {
  OOB_CUDA_OK(cudaIpcCloseMemHandle(peer_mailbox));
  return 0;
}

int oob_p2p_free(void* mailbox) {
  CtrlEntry ce{};
  if (ctrl_of(mailbox, &ce)) {
    {
      std::lock_guard<std::mutex> lk(g_ctrl_mu);
      g_ctrl.erase(mailbox);
    }
    cudaFreeHost(ce.host);
  }
  OOB_CUDA_OK(cudaFree(mailbox));
  return 0;
}

static int apply_timeout_env() {
  static bool done = false;
  if (done) return 0;
  done = true;
  if (const char* e = getenv("OOB_P2P_TIMEOUT_S")) {
    const unsigned long long ns = (unsigned long long)(atof(e) * 1e9);
    if (ns > 0) OOB_CUDA_OK(cudaMemcpyToSymbol(g_spin_timeout_ns, &ns, sizeof(ns)));
  }
  return 0;
}

/* Make every kernel waiting on this rank's mailbox give up (peer lost).  A plain store into pinned host memory: no
 * CUDA call, no stream -- callable from the listener thread while every stream of the process is blocked behind a
 * spinning kernel.  `stream` is ignored (kept for ABI compatibility with round 1). */
int oob_p2p_abort(void* mailbox, void* stream) {
  (void)stream;
  CtrlEntry ce{};
  OOB_CHECK(ctrl_of(mailbox, &ce), "oob_p2p_abort: unknown mailbox");
  ce.host->abort = 1;
  return 0;
}

/* 0 = healthy, 1 = a wait on this mailbox was aborted, 2 = a wait hit the watchdog.  Host-side read, no sync. */
int oob_p2p_status(void* mailbox, int* status) {
  CtrlEntry ce{};
  OOB_CHECK(ctrl_of(mailbox, &ce) && status, "oob_p2p_status: unknown mailbox");
  *status = (int)ce.host->status;     // set by a wait kernel that actually gave up; the bare abort request is not a fault
  return 0;
}

/* one tensor of message `seq` (1-based) -> peer ring slot.  first/last mark the first / last tensor of the message:
 * the first waits for the slot to be free, the last publishes the flag. */
int oob_p2p_send(const void* src, long bytes, void* my_mailbox, void* peer_mailbox, int nslots, long slot_bytes,
                 long offset_in_slot, unsigned seq, int first, int last, void* stream) {
  OOB_CHECK(nslots >= 1 && nslots <= P2P_MAX_SLOTS, "oob_p2p_send: nslots out of range");
  OOB_CHECK(offset_in_slot + bytes <= slot_bytes && (offset_in_slot & 15) == 0, "oob_p2p_send: payload does not fit the slot");
  OOB_CHECK((reinterpret_cast<uintptr_t>(src) & 15) == 0, "oob_p2p_send: source must be 16 B aligned");
  MailboxHeader* mine = reinterpret_cast<MailboxHeader*>(my_mailbox);
  MailboxHeader* peer = reinterpret_cast<MailboxHeader*>(peer_mailbox);
  const int slot = (int)((seq - 1) % (unsigned)nslots);
  char* dst = reinterpret_cast<char*>(peer + 1) + (long)slot * slot_bytes + offset_in_slot;
  int blocks = (int)((bytes / 16 + 256 * 4 - 1) / (256 * 4));
  blocks = blocks < 1 ? 1 : (blocks > 64 ? 64 : blocks);
  const int wait_ack = first && seq > (unsigned)nslots;
  if (int rc = apply_timeout_env()) return rc;
  CtrlEntry ce{};
  OOB_CHECK(ctrl_of(my_mailbox, &ce), "oob_p2p_send: unknown mailbox");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (wait_ack) {   // the peer must have drained this slot (message seq - nslots)
    p2p_wait_kernel<<<1, 32, 0, st>>>(&mine->acks[slot], seq - (unsigned)nslots, &mine->abort, ce.dev);
    OOB_CUDA_OK(cudaGetLastError());
    count_launch();
  }
  p2p_copy_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const char*>(src), dst, bytes,
                                          last ? &peer->flags[slot] : nullptr, seq, &mine->counters[0], &mine->abort);
  OOB_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

int oob_p2p_recv(void* dst, long bytes, void* my_mailbox, void* peer_mailbox, int nslots, long slot_bytes,
                 long offset_in_slot, unsigned seq, int first, int last, void* stream) {
  OOB_CHECK(nslots >= 1 && nslots <= P2P_MAX_SLOTS, "oob_p2p_recv: nslots out of range");
  OOB_CHECK(offset_in_slot + bytes <= slot_bytes && (offset_in_slot & 15) == 0, "oob_p2p_recv: payload does not fit the slot");
  OOB_CHECK((reinterpret_cast<uintptr_t>(dst) & 15) == 0, "oob_p2p_recv: destination must be 16 B aligned");
  MailboxHeader* mine = reinterpret_cast<MailboxHeader*>(my_mailbox);
  MailboxHeader* peer = reinterpret_cast<MailboxHeader*>(peer_mailbox);
  const int slot = (int)((seq - 1) % (unsigned)nslots);
  const char* src = reinterpret_cast<const char*>(mine + 1) + (long)slot * slot_bytes + offset_in_slot;
  int blocks = (int)((bytes / 16 + 256 * 4 - 1) / (256 * 4));
  blocks = blocks < 1 ? 1 : (blocks > 64 ? 64 : blocks);
  if (int rc = apply_timeout_env()) return rc;
  CtrlEntry ce{};
  OOB_CHECK(ctrl_of(my_mailbox, &ce), "oob_p2p_recv: unknown mailbox");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (first) {      // the peer's flag for this message
    p2p_wait_kernel<<<1, 32, 0, st>>>(&mine->flags[slot], seq, &mine->abort, ce.dev);
    OOB_CUDA_OK(cudaGetLastError());
    count_launch();
  }
  p2p_copy_kernel<<<blocks, 256, 0, st>>>(src, reinterpret_cast<char*>(dst), bytes, last ? &peer->acks[slot] : nullptr,
                                          seq, &mine->counters[1], &mine->abort);
  OOB_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

}  // extern "C"
