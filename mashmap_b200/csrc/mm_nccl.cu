/*
 * mm_nccl.cu -- libmashmap_nccl.so: the multi-GPU entry points declared in include/mashmap_b200_nccl.h, written on top
 * of the public C ABI of libmashmap_b200.so (mm_index_blob / mm_index_blob_alloc / mm_index_adopt_blob) and NCCL.
 * one broadcast of the index image over NVLink, reads sharded by rank, one all-gather of mapping records.
 */
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mashmap_b200_nccl.h"

static thread_local std::string g_comm_error;

struct mm_comm {
  ncclComm_t comm = nullptr;
  cudaStream_t stream = nullptr;
  int n_ranks = 0, rank = 0, device = 0;
  std::string error;
  /* staging, grown on demand */
  unsigned char *d_send = nullptr, *d_recv = nullptr;
  uint64_t send_cap = 0, recv_cap = 0;
  unsigned long long *d_counts = nullptr; /* [n_ranks + 1]: slot n_ranks = this rank's own count / a size to broadcast */
  unsigned long long *h_counts = nullptr; /* pinned */
};

namespace {

int cfail(mm_comm *c, int code, const char *fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->error = buf; else g_comm_error = buf;
  return code;
}

#define CUC(c, call)                                                                                          \
  do {                                                                                                        \
    cudaError_t e_ = (call);                                                                                  \
    if (e_ != cudaSuccess) return cfail(c, e_ == cudaErrorMemoryAllocation ? MM_ENOMEM : MM_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)
#define NCC(c, call)                                                                            \
  do {                                                                                          \
    ncclResult_t r_ = (call);                                                                   \
    if (r_ != ncclSuccess) return cfail(c, MM_ECUDA, "%s: %s", #call, ncclGetErrorString(r_)); \
  } while (0)

} // namespace

extern "C" {

int mm_comm_unique_id(uint8_t id[MM_COMM_ID_BYTES])
{
  static_assert(sizeof(ncclUniqueId) == MM_COMM_ID_BYTES, "NCCL unique id size");
  if (!id) return MM_EINVAL;
  ncclUniqueId u;
  NCC(nullptr, ncclGetUniqueId(&u));
  memcpy(id, &u, sizeof u);
  return MM_OK;
}

int mm_comm_create(const uint8_t id[MM_COMM_ID_BYTES], int n_ranks, int rank, int device, mm_comm **out)
{
  if (!id || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks) return cfail(nullptr, MM_EINVAL, "bad communicator arguments");
  *out = nullptr;
  CUC(nullptr, cudaSetDevice(device));
  mm_comm *c = new mm_comm();
  c->n_ranks = n_ranks; c->rank = rank; c->device = device;
  ncclUniqueId u;
  memcpy(&u, id, sizeof u);
  ncclResult_t r = ncclCommInitRank(&c->comm, n_ranks, u, rank);
  if (r != ncclSuccess) { delete c; return cfail(nullptr, MM_ECUDA, "ncclCommInitRank: %s", ncclGetErrorString(r)); }
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMalloc((void **)&c->d_counts, ((size_t)n_ranks + 1) * 8) != cudaSuccess ||
      cudaHostAlloc((void **)&c->h_counts, ((size_t)n_ranks + 1) * 8, cudaHostAllocDefault) != cudaSuccess) {
    ncclCommDestroy(c->comm);
    delete c;
    return cfail(nullptr, MM_ENOMEM, "cannot allocate communicator buffers");
  }
  *out = c;
  return MM_OK;
}

int mm_comm_destroy(mm_comm *c)
{
  if (!c) return MM_OK;
  cudaSetDevice(c->device);
  if (c->comm) ncclCommDestroy(c->comm);
  if (c->stream) cudaStreamDestroy(c->stream);
  cudaFree(c->d_send); cudaFree(c->d_recv); cudaFree(c->d_counts);
  if (c->h_counts) cudaFreeHost(c->h_counts);
  delete c;
  return MM_OK;
}

const char *mm_comm_last_error(const mm_comm *c) { return c ? c->error.c_str() : g_comm_error.c_str(); }

int mm_index_broadcast(mm_ctx *ctx, mm_comm *c, int root, uint64_t *n_bytes)
{
  if (!ctx || !c || root < 0 || root >= c->n_ranks) return cfail(c, MM_EINVAL, "bad broadcast arguments");
  CUC(c, cudaSetDevice(c->device));
  void *blob = nullptr;
  uint64_t n = 0;
  if (c->rank == root) {
    int rc = mm_index_blob(ctx, &blob, &n);
    if (rc != MM_OK) return cfail(c, rc, "root has no index image: %s", mm_last_error(ctx));
  }
  /* the size first (8 bytes), then the image: one broadcast each */
  c->h_counts[0] = n;
  CUC(c, cudaMemcpyAsync(c->d_counts, c->h_counts, 8, cudaMemcpyHostToDevice, c->stream));
  NCC(c, ncclBroadcast(c->d_counts, c->d_counts, 1, ncclUint64, root, c->comm, c->stream));
  CUC(c, cudaMemcpyAsync(c->h_counts, c->d_counts, 8, cudaMemcpyDeviceToHost, c->stream));
  CUC(c, cudaStreamSynchronize(c->stream));
  n = c->h_counts[0];
  if (n == 0) return cfail(c, MM_ESTATE, "the root broadcast an empty index image");
  if (c->rank != root) {
    int rc = mm_index_blob_alloc(ctx, n, &blob);
    if (rc != MM_OK) return cfail(c, rc, "cannot allocate the index image (%llu bytes): %s", (unsigned long long)n, mm_last_error(ctx));
  }
  NCC(c, ncclBroadcast(blob, blob, n, ncclUint8, root, c->comm, c->stream));
  /* the context reads the image on its own stream: it must be complete on the device before it is adopted */
  CUC(c, cudaStreamSynchronize(c->stream));
  if (c->rank != root) {
    int rc = mm_index_adopt_blob(ctx);
    if (rc != MM_OK) return cfail(c, rc, "adopting the broadcast image failed: %s", mm_last_error(ctx));
  }
  if (n_bytes) *n_bytes = n;
  return MM_OK;
}

int mm_records_allgather(mm_comm *c, const void *records, uint64_t n_records, uint32_t record_bytes, void *out,
                         uint64_t out_cap_records, uint64_t *counts)
{
  if (!c || !counts || record_bytes == 0 || (!records && n_records)) return cfail(c, MM_EINVAL, "bad all-gather arguments");
  CUC(c, cudaSetDevice(c->device));
  const int N = c->n_ranks;
  /* counts */
  c->h_counts[N] = n_records;
  CUC(c, cudaMemcpyAsync(c->d_counts + N, c->h_counts + N, 8, cudaMemcpyHostToDevice, c->stream));
  NCC(c, ncclAllGather(c->d_counts + N, c->d_counts, 1, ncclUint64, c->comm, c->stream));
  CUC(c, cudaMemcpyAsync(c->h_counts, c->d_counts, (size_t)N * 8, cudaMemcpyDeviceToHost, c->stream));
  CUC(c, cudaStreamSynchronize(c->stream));
  uint64_t m = 1, total = 0;
  for (int r = 0; r < N; r++) { counts[r] = c->h_counts[r]; total += counts[r]; if (counts[r] > m) m = counts[r]; }
  if (total > out_cap_records || (!out && total)) return cfail(c, MM_ECAPACITY, "need room for %llu records", (unsigned long long)total);
  /* padded records: every rank contributes m records */
  const uint64_t slot = m * (uint64_t)record_bytes;
  if (c->send_cap < slot) {
    cudaFree(c->d_send); c->d_send = nullptr; c->send_cap = 0;
    CUC(c, cudaMalloc((void **)&c->d_send, slot + slot / 4));
    c->send_cap = slot + slot / 4;
  }
  if (c->recv_cap < slot * (uint64_t)N) {
    cudaFree(c->d_recv); c->d_recv = nullptr; c->recv_cap = 0;
    CUC(c, cudaMalloc((void **)&c->d_recv, (slot + slot / 4) * (uint64_t)N));
    c->recv_cap = (slot + slot / 4) * (uint64_t)N;
  }
  if (n_records) CUC(c, cudaMemcpyAsync(c->d_send, records, n_records * (uint64_t)record_bytes, cudaMemcpyHostToDevice, c->stream));
  NCC(c, ncclAllGather(c->d_send, c->d_recv, slot, ncclUint8, c->comm, c->stream));
  unsigned char *o = (unsigned char *)out;
  for (int r = 0; r < N; r++) {
    if (counts[r]) CUC(c, cudaMemcpyAsync(o, c->d_recv + (uint64_t)r * slot, counts[r] * (uint64_t)record_bytes, cudaMemcpyDeviceToHost, c->stream));
    o += counts[r] * (uint64_t)record_bytes;
  }
  CUC(c, cudaStreamSynchronize(c->stream));
  return MM_OK;
}

int mm_index_replicate(mm_ctx *src, mm_ctx *const *dst, int n_dst)
{
  if (!src || (!dst && n_dst) || n_dst < 0) return cfail(nullptr, MM_EINVAL, "bad replicate arguments");
  if (n_dst == 0) return MM_OK;
  const int N = n_dst + 1;
  std::vector<int> devs(N);
  std::vector<mm_ctx *> ctxs(N);
  ctxs[0] = src;
  for (int i = 0; i < n_dst; i++) ctxs[i + 1] = dst[i];
  for (int i = 0; i < N; i++) {
    devs[i] = mm_ctx_device(ctxs[i]);
    for (int j = 0; j < i; j++)
      if (devs[j] == devs[i]) return cfail(nullptr, MM_EINVAL, "two contexts on device %d: share the image with mm_ctx_share_index instead", devs[i]);
  }
  void *blob0 = nullptr;
  uint64_t n = 0;
  int rc = mm_index_blob(src, &blob0, &n);
  if (rc != MM_OK) return cfail(nullptr, rc, "source context has no index image: %s", mm_last_error(src));
  std::vector<void *> blobs(N, nullptr);
  blobs[0] = blob0;
  for (int i = 1; i < N; i++)
    if ((rc = mm_index_blob_alloc(ctxs[i], n, &blobs[i])) != MM_OK)
      return cfail(nullptr, rc, "device %d: cannot allocate the index image: %s", devs[i], mm_last_error(ctxs[i]));
  std::vector<ncclComm_t> comms(N);
  NCC(nullptr, ncclCommInitAll(comms.data(), N, devs.data()));
  std::vector<cudaStream_t> streams(N);
  for (int i = 0; i < N; i++) {
    CUC(nullptr, cudaSetDevice(devs[i]));
    CUC(nullptr, cudaStreamCreateWithFlags(&streams[i], cudaStreamNonBlocking));
  }
  NCC(nullptr, ncclGroupStart());
  for (int i = 0; i < N; i++) NCC(nullptr, ncclBroadcast(blobs[i], blobs[i], n, ncclUint8, 0, comms[i], streams[i]));
  NCC(nullptr, ncclGroupEnd());
  for (int i = 0; i < N; i++) {
    CUC(nullptr, cudaSetDevice(devs[i]));
    CUC(nullptr, cudaStreamSynchronize(streams[i]));
    cudaStreamDestroy(streams[i]);
    ncclCommDestroy(comms[i]);
  }
  for (int i = 1; i < N; i++)
    if ((rc = mm_index_adopt_blob(ctxs[i])) != MM_OK)
      return cfail(nullptr, rc, "device %d: adopting the image failed: %s", devs[i], mm_last_error(ctxs[i]));
  return MM_OK;
}

} // extern "C"
