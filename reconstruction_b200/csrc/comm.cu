// comm.cu — the one exchange step of the path behind the C ABI: after DisparityToCloud every
// rank holds the points of its own camera pair(s); the sink wants all of them in pair order
// (CloudOptimization/CCloudOptimization.cpp:123 appends the clouds pair by pair; the reference loops the pairs serially,
// CStereoMatching.cpp:17-33).  One NCCL communicator per GPU (one process per GPU, or one host thread per GPU inside a
// process), and per exchanged pair: an all-gather of the counts, then ONE NCCL group in which every rank sends exactly
// its own points (xyz f64 x3, bgr u8 x3, pixel index i32) to every other rank's slice of the gathered buffers —
// nothing is padded to the largest count.
//
// The matcher never waits for another rank: a context only SNAPSHOTS its points (device copy on its own stream into a
// staging slot) and returns; one exchange thread per communicator issues the collectives in ticket order (the same
// order on every rank) on a dedicated highest-priority stream, with NCCL's CTA count bounded (SB200_NCCL_MAX_CTAS): the
// collective's kernels take few SMs from the matcher, but get them at once.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): the library loads and every other entry point works on a machine
// without NCCL; sb200_comm_* then report the missing library.
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <condition_variable>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/stereo_b200.h"
#include "common.cuh"

namespace {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  std::string err;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, []() {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { api.err = std::string("NCCL not found: ") + dlerror(); return; }
#define SB_SYM(field, name)                                               \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name));      \
  if (!api.field) { api.err = std::string("NCCL lacks ") + name; return; }
    SB_SYM(GetUniqueId, "ncclGetUniqueId")
    SB_SYM(CommInitRank, "ncclCommInitRank")
    SB_SYM(CommDestroy, "ncclCommDestroy")
    SB_SYM(AllGather, "ncclAllGather")
    SB_SYM(Broadcast, "ncclBroadcast")
    SB_SYM(Send, "ncclSend")
    SB_SYM(Recv, "ncclRecv")
    SB_SYM(GroupStart, "ncclGroupStart")
    SB_SYM(GroupEnd, "ncclGroupEnd")
    SB_SYM(GetErrorString, "ncclGetErrorString")
#undef SB_SYM
    api.CommInitRankConfig = reinterpret_cast<decltype(api.CommInitRankConfig)>(dlsym(h, "ncclCommInitRankConfig"));  // optional
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(dlsym(h, "ncclGetVersion"));
    api.ok = true;
  });
  return api;
}

struct Job {
  int producer = 0, slot = 0;
  int64_t n = 0;
  cudaEvent_t ready = nullptr;  // snapshot copy done (recorded on the producer's stream)
};

struct Stage {  // one staging slot of one producer
  double* xyz = nullptr;
  uint8_t* bgr = nullptr;
  int* pix = nullptr;
  int64_t cap = 0;
  bool busy = false;            // snapshot taken, gather not yet enqueued
  cudaEvent_t read_done = nullptr;  // the gather that read this slot has finished (recorded on the exchange stream)
  bool has_read = false;
};

struct Gathered {  // result of one ticket
  double* xyz = nullptr;
  uint8_t* bgr = nullptr;
  int* pix = nullptr;
  int64_t cap = 0, total = 0;
  std::vector<int64_t> counts;
  cudaEvent_t t0 = nullptr, t1 = nullptr, done = nullptr;
  int64_t ticket = -1;
  bool timed = false;
};

}  // namespace

struct sb200_comm {
  int device = 0, rank = 0, nranks = 1, producers = 1, slots = 2;
  ncclComm_t comm = nullptr;
  cudaStream_t xs = nullptr;  // exchange stream
  cudaEvent_t ev_counts = nullptr;  // blocking-sync event: the exchange thread sleeps while it waits for the counts
  std::vector<Stage> stage;   // [producer * slots + slot]
  Gathered out[2];
  long long* d_counts = nullptr;  // [nranks + 1]: gathered counts, then this rank's count
  long long* h_counts = nullptr;  // pinned [nranks]
  std::mutex mu;
  std::condition_variable cv;
  std::map<int64_t, Job> queue;
  int64_t next_ticket = 0;    // next ticket the exchange thread will gather
  int64_t gathered = 0;       // tickets [0, gathered) have been enqueued on the exchange stream
  bool stop = false;
  std::thread worker;
  std::string err;
  bool failed = false;
  double coll_ms = 0;
  int64_t coll_bytes = 0, coll_n = 0;
  int64_t sync_seq = 0;       // sequence counter of sb200_allgather_points
  // all-gather mode: this rank's block padded to the largest count (send side) and the padded gathered blocks (receive side)
  double* sx = nullptr; uint8_t* sbg = nullptr; int* spx = nullptr; int64_t scap = 0;
  double* px = nullptr; uint8_t* pbg = nullptr; int* ppx = nullptr; int64_t pcap = 0;
  bool track_consumer = false;  // the exchange thread reuses a result slot only after sb200_exchange_wait returned for the ticket in it
  int64_t consumed = 0;       // tickets [0, consumed) have been waited for
};

namespace {

#define CKC(call)                                                                                          \
  do {                                                                                                     \
    cudaError_t e_ = (call);                                                                               \
    if (e_ != cudaSuccess) {                                                                               \
      char b_[512];                                                                                        \
      snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      c->err = b_;                                                                                         \
      return SB200_ERR_CUDA;                                                                               \
    }                                                                                                      \
  } while (0)
#define CKN(call)                                                                                                   \
  do {                                                                                                              \
    ncclResult_t r_ = (call);                                                                                       \
    if (r_ != ncclSuccess) {                                                                                        \
      char b_[512];                                                                                                 \
      snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, nccl().GetErrorString(r_), __FILE__, __LINE__);       \
      c->err = b_;                                                                                                  \
      return SB200_ERR_CUDA;                                                                                        \
    }                                                                                                               \
  } while (0)

int grow(sb200_comm* c, double** xyz, uint8_t** bgr, int** pix, int64_t* cap, int64_t need) {
  if (need <= *cap) return SB200_OK;
  const int64_t n = need + need / 8 + 1024;
  cudaFree(*xyz); cudaFree(*bgr); cudaFree(*pix);
  *xyz = nullptr; *bgr = nullptr; *pix = nullptr; *cap = 0;
  CKC(cudaMalloc((void**)xyz, (size_t)n * 24));
  CKC(cudaMalloc((void**)bgr, (size_t)n * 3));
  CKC(cudaMalloc((void**)pix, (size_t)n * 4));
  *cap = n;
  return SB200_OK;
}

// the collectives of one ticket, on the exchange stream (exchange thread only)
int gather_one(sb200_comm* c, int64_t ticket, const Job& job) {
  NcclApi& N = nccl();
  Stage& st = c->stage[(size_t)job.producer * c->slots + job.slot];
  Gathered& g = c->out[ticket & 1];
  if (g.ticket >= 0 && !g.timed) {  // the previous occupant (ticket - 2) finished long ago: book its device time before the events are reused
    if (cudaEventSynchronize(g.t1) == cudaSuccess) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, g.t0, g.t1) == cudaSuccess) {
        std::lock_guard<std::mutex> lk(c->mu);
        c->coll_ms += ms;
      }
    }
    (void)cudaGetLastError();
    g.timed = true;
  }
  if (job.ready) {
    CKC(cudaStreamWaitEvent(c->xs, job.ready, 0));
    cudaEventDestroy(job.ready);  // the wait is enqueued; the event's resources go when it has completed
  }
  // the previous result in this output slot (ticket - 2) must have been consumed: its `done` event precedes us on the same stream
  CKC(cudaEventRecord(g.t0, c->xs));
  const long long mine = job.n;
  CKC(cudaMemcpyAsync(c->d_counts + c->nranks, &mine, sizeof mine, cudaMemcpyHostToDevice, c->xs));
  CKN(N.AllGather(c->d_counts + c->nranks, c->d_counts, 1, ncclInt64, c->comm, c->xs));
  CKC(cudaMemcpyAsync(c->h_counts, c->d_counts, sizeof(long long) * c->nranks, cudaMemcpyDeviceToHost, c->xs));
  CKC(cudaEventRecord(c->ev_counts, c->xs));
  CKC(cudaEventSynchronize(c->ev_counts));  // the only host-blocking step; it blocks this thread alone, and sleeping (blocking-sync event)
  g.counts.assign(c->nranks, 0);
  int64_t total = 0;
  for (int r = 0; r < c->nranks; r++) { g.counts[r] = c->h_counts[r]; total += c->h_counts[r]; }
  int rc = grow(c, &g.xyz, &g.bgr, &g.pix, &g.cap, total);
  if (rc) return rc;
  g.total = total;
  // Every rank's block goes to every other rank, un-padded, in ONE group.  Default: direct sends and receives (through the
  // NVSwitch every pair of GPUs has its own full-bandwidth path, so the 7 + 7 transfers of a rank run side by side; measured at
  // 8 GPUs, 5.4 GB received per rank and step).  SB200_EXCHANGE_MODE=bcast: one ncclBroadcast per rank and array instead - ring
  // broadcasts, which at 8 ranks kept the exchange stream busy for longer than a step takes (103 ms per 93 ms step, 0.83 scaling).
  // mode: "sendrecv" (default) = one group of direct sends / receives, nothing padded; "allgather" = three ncclAllGather on blocks
  // padded to the largest count, then device copies that close the gaps (the gathered buffers stay un-padded, rank-major);
  // "bcast" = one ncclBroadcast per rank and array.  Measured at 8 GPUs, 5.4 GB received per rank and step (config C, three pairs
  // in flight, 16 CTAs): sendrecv 2950, allgather 2878, bcast 2727 (8 CTAs), NCCL's copy-engine p2p (NCCL_P2P_USE_CUDA_MEMCPY=1)
  // 2658 Mpix/s; without the exchange 3234 (profiles/r2_bench_n8_*.json).
  static const int mode = [] {
    const char* e = getenv("SB200_EXCHANGE_MODE");
    return !e ? 1 : !strcmp(e, "allgather") ? 0 : !strcmp(e, "bcast") ? 2 : 1;
  }();
  std::vector<int64_t> offs(c->nranks, 0);
  int64_t nmax = g.counts[0];
  for (int r = 1; r < c->nranks; r++) { offs[r] = offs[r - 1] + g.counts[r - 1]; nmax = nmax > g.counts[r] ? nmax : g.counts[r]; }
  if (mode == 0 && c->nranks > 1 && nmax > 0) {
    rc = grow(c, &c->sx, &c->sbg, &c->spx, &c->scap, nmax);
    if (rc) return rc;
    rc = grow(c, &c->px, &c->pbg, &c->ppx, &c->pcap, nmax * c->nranks);
    if (rc) return rc;
    const int64_t n = g.counts[c->rank];
    if (n > 0) {
      CKC(cudaMemcpyAsync(c->sx, st.xyz, (size_t)n * 24, cudaMemcpyDeviceToDevice, c->xs));
      CKC(cudaMemcpyAsync(c->sbg, st.bgr, (size_t)n * 3, cudaMemcpyDeviceToDevice, c->xs));
      CKC(cudaMemcpyAsync(c->spx, st.pix, (size_t)n * 4, cudaMemcpyDeviceToDevice, c->xs));
    }
    CKN(N.GroupStart());
    CKN(N.AllGather(c->sx, c->px, (size_t)nmax * 3, ncclDouble, c->comm, c->xs));
    CKN(N.AllGather(c->sbg, c->pbg, (size_t)nmax * 3, ncclUint8, c->comm, c->xs));
    CKN(N.AllGather(c->spx, c->ppx, (size_t)nmax, ncclInt32, c->comm, c->xs));
    CKN(N.GroupEnd());
    for (int r = 0; r < c->nranks; r++) {  // close the gaps: rank r's rows to their place in the rank-major concatenation
      const int64_t nr = g.counts[r], off = offs[r];
      if (nr <= 0) continue;
      CKC(cudaMemcpyAsync(g.xyz + 3 * off, c->px + 3 * r * nmax, (size_t)nr * 24, cudaMemcpyDeviceToDevice, c->xs));
      CKC(cudaMemcpyAsync(g.bgr + 3 * off, c->pbg + 3 * r * nmax, (size_t)nr * 3, cudaMemcpyDeviceToDevice, c->xs));
      CKC(cudaMemcpyAsync(g.pix + off, c->ppx + r * nmax, (size_t)nr * 4, cudaMemcpyDeviceToDevice, c->xs));
    }
  } else if (mode == 2) {
    CKN(N.GroupStart());
    for (int r = 0; r < c->nranks; r++) {
      const int64_t n = g.counts[r], off = offs[r];
      if (n <= 0) continue;
      const bool me = r == c->rank;
      CKN(N.Broadcast(me ? (const void*)st.xyz : (const void*)(g.xyz + 3 * off), g.xyz + 3 * off, (size_t)n * 3, ncclDouble, r, c->comm, c->xs));
      CKN(N.Broadcast(me ? (const void*)st.bgr : (const void*)(g.bgr + 3 * off), g.bgr + 3 * off, (size_t)n * 3, ncclUint8, r, c->comm, c->xs));
      CKN(N.Broadcast(me ? (const void*)st.pix : (const void*)(g.pix + off), g.pix + off, (size_t)n, ncclInt32, r, c->comm, c->xs));
    }
    CKN(N.GroupEnd());
  } else {
    const int64_t mine_n = g.counts[c->rank];
    CKN(N.GroupStart());
    for (int d = 1; d < c->nranks; d++) {  // peers in a rotated order: rank r talks to r+d and r-d in step d
      const int to = (c->rank + d) % c->nranks, from = (c->rank - d + c->nranks) % c->nranks;
      if (mine_n > 0) {
        CKN(N.Send(st.xyz, (size_t)mine_n * 3, ncclDouble, to, c->comm, c->xs));
        CKN(N.Send(st.bgr, (size_t)mine_n * 3, ncclUint8, to, c->comm, c->xs));
        CKN(N.Send(st.pix, (size_t)mine_n, ncclInt32, to, c->comm, c->xs));
      }
      const int64_t n = g.counts[from], off = offs[from];
      if (n > 0) {
        CKN(N.Recv(g.xyz + 3 * off, (size_t)n * 3, ncclDouble, from, c->comm, c->xs));
        CKN(N.Recv(g.bgr + 3 * off, (size_t)n * 3, ncclUint8, from, c->comm, c->xs));
        CKN(N.Recv(g.pix + off, (size_t)n, ncclInt32, from, c->comm, c->xs));
      }
    }
    CKN(N.GroupEnd());
    if (mine_n > 0) {  // this rank's own block
      const int64_t off = offs[c->rank];
      CKC(cudaMemcpyAsync(g.xyz + 3 * off, st.xyz, (size_t)mine_n * 24, cudaMemcpyDeviceToDevice, c->xs));
      CKC(cudaMemcpyAsync(g.bgr + 3 * off, st.bgr, (size_t)mine_n * 3, cudaMemcpyDeviceToDevice, c->xs));
      CKC(cudaMemcpyAsync(g.pix + off, st.pix, (size_t)mine_n * 4, cudaMemcpyDeviceToDevice, c->xs));
    }
  }
  CKC(cudaEventRecord(g.t1, c->xs));
  CKC(cudaEventRecord(g.done, c->xs));
  CKC(cudaEventRecord(st.read_done, c->xs));
  g.ticket = ticket;
  g.timed = false;
  c->coll_bytes += total * 31;
  c->coll_n++;
  return SB200_OK;
}

void exchange_thread(sb200_comm* c) {
  cudaSetDevice(c->device);
  for (;;) {
    Job job;
    int64_t ticket;
    {
      std::unique_lock<std::mutex> lk(c->mu);
      c->cv.wait(lk, [&]() { return c->stop || c->queue.count(c->next_ticket) != 0; });
      if (c->queue.count(c->next_ticket) == 0) return;  // stop requested and nothing left in order
      ticket = c->next_ticket;
      // two results are kept (ticket parity): with a consumer registered, ticket t may only overwrite ticket t-2 once that one
      // has been handed over (sb200_exchange_wait returned) - flow control towards a consumer slower than the exchange
      if (c->track_consumer) {
        c->cv.wait(lk, [&]() { return c->stop || c->consumed >= ticket - 1; });
        if (c->stop && c->consumed < ticket - 1) return;
      }
      job = c->queue[ticket];
      c->queue.erase(ticket);
    }
    const int rc = gather_one(c, ticket, job);
    {
      std::lock_guard<std::mutex> lk(c->mu);
      Stage& st = c->stage[(size_t)job.producer * c->slots + job.slot];
      st.busy = false;
      st.has_read = rc == SB200_OK;
      if (rc != SB200_OK) c->failed = true;
      c->next_ticket = ticket + 1;
      c->gathered = ticket + 1;
    }
    c->cv.notify_all();
    if (rc != SB200_OK) return;
  }
}

void add_time(sb200_comm* c, Gathered& g) {
  if (g.timed || g.ticket < 0) return;
  float ms = 0;
  if (cudaEventElapsedTime(&ms, g.t0, g.t1) == cudaSuccess) {
    std::lock_guard<std::mutex> lk(c->mu);
    c->coll_ms += ms;
  } else {
    (void)cudaGetLastError();
  }
  g.timed = true;
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

int sb200_comm_unique_id(void* id_out) {
  if (!id_out) return SB200_ERR_BAD_ARG;
  NcclApi& N = nccl();
  if (!N.ok) return SB200_ERR_NO_DEVICE;
  ncclUniqueId id;
  if (N.GetUniqueId(&id) != ncclSuccess) return SB200_ERR_CUDA;
  static_assert(sizeof(ncclUniqueId) == SB200_UNIQUE_ID_BYTES, "ncclUniqueId size");
  memcpy(id_out, &id, sizeof id);
  return SB200_OK;
}

const char* sb200_comm_last_error(const sb200_comm* c) {
  if (!c) return nccl().err.c_str();
  return c->err.c_str();
}

int sb200_comm_init(sb200_comm** out, int device, int rank, int nranks, const void* unique_id, int producers, int slots) {
  if (!out) return SB200_ERR_BAD_ARG;
  *out = nullptr;
  if (!unique_id || nranks < 1 || rank < 0 || rank >= nranks || producers < 1 || slots < 1) return SB200_ERR_BAD_ARG;
  NcclApi& N = nccl();
  if (!N.ok) return SB200_ERR_NO_DEVICE;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return SB200_ERR_NO_DEVICE;
  sb200_comm* c = new sb200_comm();
  *out = c;  // returned even on failure so the caller can read sb200_comm_last_error, then destroy
  c->device = device; c->rank = rank; c->nranks = nranks; c->producers = producers; c->slots = slots;
  CKC(cudaSetDevice(device));
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof id);
  int max_ctas = 16;  // a bounded share of the SMs for the collective (measured on two GPUs: 4 -> 766, 8 -> 786, 16 -> 795 Mpix/s)
  if (const char* e = getenv("SB200_NCCL_MAX_CTAS")) max_ctas = atoi(e);
  if (N.CommInitRankConfig && max_ctas > 0) {
    ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
    cfg.maxCTAs = max_ctas;
    cfg.minCTAs = 1;
    CKN(N.CommInitRankConfig(&c->comm, nranks, id, rank, &cfg));
  } else {
    CKN(N.CommInitRank(&c->comm, nranks, id, rank));
  }
  // Highest priority: the collective's few CTAs (SB200_NCCL_MAX_CTAS) must get SM slots as soon as they are launched.  The
  // matcher keeps every SM saturated with thousands of queued CTAs; on a LOW-priority stream NCCL's CTAs were only scheduled when
  // those queues ran dry, the exchange fell behind the producers (two staging slots) and an 8-GPU step took 183 ms instead of 100.
  int lo = 0, hi = 0;
  CKC(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // hi = numerically smallest = highest priority
  CKC(cudaStreamCreateWithPriority(&c->xs, cudaStreamNonBlocking, hi));
  CKC(cudaEventCreateWithFlags(&c->ev_counts, cudaEventBlockingSync | cudaEventDisableTiming));
  c->stage.resize((size_t)producers * slots);
  for (Stage& s : c->stage) CKC(cudaEventCreateWithFlags(&s.read_done, cudaEventDisableTiming));
  for (Gathered& g : c->out) {
    CKC(cudaEventCreate(&g.t0));
    CKC(cudaEventCreate(&g.t1));
    CKC(cudaEventCreateWithFlags(&g.done, cudaEventBlockingSync | cudaEventDisableTiming));
  }
  CKC(cudaMalloc((void**)&c->d_counts, sizeof(long long) * (nranks + 1)));
  CKC(cudaMallocHost((void**)&c->h_counts, sizeof(long long) * nranks));
  c->worker = std::thread(exchange_thread, c);
  return SB200_OK;
}

void sb200_comm_destroy(sb200_comm* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  {
    std::lock_guard<std::mutex> lk(c->mu);
    c->stop = true;
  }
  c->cv.notify_all();
  if (c->worker.joinable()) c->worker.join();
  if (c->xs) cudaStreamSynchronize(c->xs);
  for (auto& kv : c->queue) if (kv.second.ready) cudaEventDestroy(kv.second.ready);
  for (Stage& s : c->stage) { cudaFree(s.xyz); cudaFree(s.bgr); cudaFree(s.pix); if (s.read_done) cudaEventDestroy(s.read_done); }
  for (Gathered& g : c->out) {
    cudaFree(g.xyz); cudaFree(g.bgr); cudaFree(g.pix);
    if (g.t0) cudaEventDestroy(g.t0);
    if (g.t1) cudaEventDestroy(g.t1);
    if (g.done) cudaEventDestroy(g.done);
  }
  if (c->ev_counts) cudaEventDestroy(c->ev_counts);
  cudaFree(c->sx); cudaFree(c->sbg); cudaFree(c->spx); cudaFree(c->px); cudaFree(c->pbg); cudaFree(c->ppx);
  cudaFree(c->d_counts);
  if (c->h_counts) cudaFreeHost(c->h_counts);
  if (c->comm) nccl().CommDestroy(c->comm);
  if (c->xs) cudaStreamDestroy(c->xs);
  delete c;
}

int sb200_exchange_submit(sb200_comm* c, sb200_ctx* ctx, int producer, int64_t seq) {
  if (!c || producer < 0 || producer >= c->producers || seq < 0) return SB200_ERR_BAD_ARG;
  void *xyz = nullptr, *bgr = nullptr, *pix = nullptr;
  int64_t n = 0;
  cudaStream_t ps = nullptr;
  if (ctx) {  // ctx == NULL: this rank has no pair for this ticket and contributes no points (every rank must take part)
    int rc = sb200_points_device(ctx, &xyz, &bgr, &pix, &n);
    if (rc) return rc;
    ps = (cudaStream_t)sb200_stream(ctx);
  }
  CKC(cudaSetDevice(c->device));
  const int slot = (int)(seq % c->slots);
  Stage& st = c->stage[(size_t)producer * c->slots + slot];
  {
    std::unique_lock<std::mutex> lk(c->mu);
    c->cv.wait(lk, [&]() { return !st.busy || c->failed; });  // at most `slots` snapshots of a producer wait for the exchange
    if (c->failed) return SB200_ERR_CUDA;
    st.busy = true;
  }
  Job job;
  job.producer = producer; job.slot = slot; job.n = n;
  if (n > 0) {
    if (st.has_read) CKC(cudaStreamWaitEvent(ps, st.read_done, 0));  // the gather that read this slot last must be over
    int rc = grow(c, &st.xyz, &st.bgr, &st.pix, &st.cap, n);
    if (rc) return rc;
    CKC(cudaMemcpyAsync(st.xyz, xyz, (size_t)n * 24, cudaMemcpyDeviceToDevice, ps));
    CKC(cudaMemcpyAsync(st.bgr, bgr, (size_t)n * 3, cudaMemcpyDeviceToDevice, ps));
    CKC(cudaMemcpyAsync(st.pix, pix, (size_t)n * 4, cudaMemcpyDeviceToDevice, ps));
    CKC(cudaEventCreateWithFlags(&job.ready, cudaEventDisableTiming));
    CKC(cudaEventRecord(job.ready, ps));
  }
  {
    std::lock_guard<std::mutex> lk(c->mu);
    c->queue[seq * c->producers + producer] = job;
  }
  c->cv.notify_all();
  return SB200_OK;
}

int sb200_exchange_wait(sb200_comm* c, int64_t ticket, int64_t* counts_out, double* xyz_host, uint8_t* bgr_host, int32_t* pix_host,
                        int64_t capacity, int64_t* total_out) {
  if (!c || ticket < 0) return SB200_ERR_BAD_ARG;
  {
    std::unique_lock<std::mutex> lk(c->mu);
    c->cv.wait(lk, [&]() { return c->gathered > ticket || c->failed; });
    if (c->failed) return SB200_ERR_CUDA;
    if (c->gathered > ticket + 2) { c->err = "result of this ticket was overwritten (two results are kept)"; return SB200_ERR_STATE; }
  }
  CKC(cudaSetDevice(c->device));
  Gathered& g = c->out[ticket & 1];
  CKC(cudaEventSynchronize(g.done));
  if (g.ticket != ticket) { c->err = "result of this ticket was overwritten (two results are kept)"; return SB200_ERR_STATE; }
  add_time(c, g);
  if (counts_out) for (int r = 0; r < c->nranks; r++) counts_out[r] = g.counts[r];
  if (total_out) *total_out = g.total;
  if (xyz_host || bgr_host || pix_host) {
    if (g.total > capacity) { c->err = "host buffers too small for the gathered points"; return SB200_ERR_BAD_ARG; }
    // a separate blocking copy: the exchange stream may already be busy with the next ticket
    if (xyz_host) CKC(cudaMemcpy(xyz_host, g.xyz, (size_t)g.total * 24, cudaMemcpyDeviceToHost));
    if (bgr_host) CKC(cudaMemcpy(bgr_host, g.bgr, (size_t)g.total * 3, cudaMemcpyDeviceToHost));
    if (pix_host) CKC(cudaMemcpy(pix_host, g.pix, (size_t)g.total * 4, cudaMemcpyDeviceToHost));
  }
  {
    std::lock_guard<std::mutex> lk(c->mu);
    if (ticket + 1 > c->consumed) c->consumed = ticket + 1;
  }
  c->cv.notify_all();
  return SB200_OK;
}

int sb200_comm_set_consumer(sb200_comm* c, int enable) {
  if (!c) return SB200_ERR_BAD_ARG;
  {
    std::lock_guard<std::mutex> lk(c->mu);
    c->track_consumer = enable != 0;
  }
  c->cv.notify_all();
  return SB200_OK;
}

int sb200_exchange_device(sb200_comm* c, int64_t ticket, void** xyz_dev, void** bgr_dev, void** pix_dev, int64_t* total) {
  if (!c || ticket < 0) return SB200_ERR_BAD_ARG;
  Gathered& g = c->out[ticket & 1];
  if (g.ticket != ticket) { c->err = "ticket not gathered (or already overwritten)"; return SB200_ERR_STATE; }
  if (xyz_dev) *xyz_dev = g.xyz;
  if (bgr_dev) *bgr_dev = g.bgr;
  if (pix_dev) *pix_dev = g.pix;
  if (total) *total = g.total;
  return SB200_OK;
}

int sb200_exchange_drain(sb200_comm* c, int64_t n_tickets) {
  if (!c) return SB200_ERR_BAD_ARG;
  {
    std::unique_lock<std::mutex> lk(c->mu);
    c->cv.wait(lk, [&]() { return c->gathered >= n_tickets || c->failed; });
    if (c->failed) return SB200_ERR_CUDA;
  }
  CKC(cudaSetDevice(c->device));
  CKC(cudaStreamSynchronize(c->xs));
  for (Gathered& g : c->out) add_time(c, g);
  return SB200_OK;
}

int sb200_allgather_points(sb200_comm* c, sb200_ctx* ctx, int64_t* counts_out, double* xyz_host, uint8_t* bgr_host, int32_t* pix_host,
                           int64_t capacity, int64_t* total_out) {
  if (!c || !ctx) return SB200_ERR_BAD_ARG;
  if (c->producers != 1) { c->err = "sb200_allgather_points needs a communicator with one producer"; return SB200_ERR_STATE; }
  const int64_t seq = c->sync_seq++;
  int rc = sb200_exchange_submit(c, ctx, 0, seq);
  if (rc) return rc;
  return sb200_exchange_wait(c, seq, counts_out, xyz_host, bgr_host, pix_host, capacity, total_out);
}

int sb200_comm_stats(sb200_comm* c, double* collective_ms, int64_t* bytes_received, int64_t* n_exchanges, int reset) {
  if (!c) return SB200_ERR_BAD_ARG;
  if (collective_ms) *collective_ms = c->coll_ms;
  if (bytes_received) *bytes_received = c->coll_bytes;
  if (n_exchanges) *n_exchanges = c->coll_n;
  if (reset) { c->coll_ms = 0; c->coll_bytes = 0; c->coll_n = 0; }
  return SB200_OK;
}

}  // extern "C"
#pragma GCC visibility pop
