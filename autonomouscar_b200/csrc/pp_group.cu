// pp_group.cu -- single-process multi-GPU front end.
//
// One pp_handle per device; the grid is replicated, pose batches and plan queries are
// sharded contiguously by index with NO data-path collective; NCCL (over NVLink 5 /
// NVSwitch) is used only to gather results: one ncclAllGather of the 1 B/pose flags, or of
// fixed-stride plan records, after which every device holds the whole result in input
// order.  NCCL is resolved with dlopen at pp_group_create, so the library itself loads on
// machines without NCCL (the per-GPU entry points never need it).
//
// The reference has no counterpart (single-threaded ROS node, local_planner_node.cpp:14-17);
// the multi-process variant of the same sharding (one rank per GPU under torchrun) lives in
// autonomouscar_b200/sharding.py and is what bench.py uses.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "pp_common.cuh"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;
constexpr int kNcclUint8 = 1;  // ncclDataType_t::ncclUint8

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load() {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) return false;
    CommInitAll = reinterpret_cast<decltype(CommInitAll)>(dlsym(lib, "ncclCommInitAll"));
    CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
    AllGather = reinterpret_cast<decltype(AllGather)>(dlsym(lib, "ncclAllGather"));
    GroupStart = reinterpret_cast<decltype(GroupStart)>(dlsym(lib, "ncclGroupStart"));
    GroupEnd = reinterpret_cast<decltype(GroupEnd)>(dlsym(lib, "ncclGroupEnd"));
    GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
    return CommInitAll && CommDestroy && AllGather && GroupStart && GroupEnd && GetErrorString;
  }
};

}  // namespace

struct pp_group {
  int n = 0;
  std::vector<int> devices;
  std::vector<pp_handle*> h;
  std::vector<ncclComm_t> comms;
  NcclApi nccl;
  // gather buffers, one pair per device
  std::vector<uint8_t*> d_send, d_all;
  std::vector<size_t> send_cap, all_cap;
};

#define PP_NCCL(g, expr)                                                                              \
  do {                                                                                                \
    ncclResult_t r_ = (expr);                                                                         \
    if (r_ != 0) {                                                                                    \
      ::pp::set_last_error(std::string(#expr) + ": " + (g)->nccl.GetErrorString(r_));                 \
      return PP_ERR_NCCL;                                                                             \
    }                                                                                                 \
  } while (0)

static int ensure_gather(pp_group* g, size_t per_rank_bytes) {
  for (int d = 0; d < g->n; ++d) {
    PP_CUDA(cudaSetDevice(g->devices[d]));
    if (per_rank_bytes > g->send_cap[d]) {
      cudaFree(g->d_send[d]);
      g->d_send[d] = nullptr; g->send_cap[d] = 0;
      PP_CUDA(cudaMalloc(&g->d_send[d], per_rank_bytes));
      g->send_cap[d] = per_rank_bytes;
    }
    if (per_rank_bytes * g->n > g->all_cap[d]) {
      cudaFree(g->d_all[d]);
      g->d_all[d] = nullptr; g->all_cap[d] = 0;
      PP_CUDA(cudaMalloc(&g->d_all[d], per_rank_bytes * g->n));
      g->all_cap[d] = per_rank_bytes * g->n;
    }
  }
  return PP_OK;
}

static int all_gather(pp_group* g, size_t per_rank_bytes) {
  PP_NCCL(g, g->nccl.GroupStart());
  for (int d = 0; d < g->n; ++d) {
    ncclResult_t r = g->nccl.AllGather(g->d_send[d], g->d_all[d], per_rank_bytes, kNcclUint8, g->comms[d],
                                       g->h[d]->stream);
    if (r != 0) {
      g->nccl.GroupEnd();
      pp::set_last_error(std::string("ncclAllGather: ") + g->nccl.GetErrorString(r));
      return PP_ERR_NCCL;
    }
  }
  PP_NCCL(g, g->nccl.GroupEnd());
  return PP_OK;
}

extern "C" {

int pp_group_create(const pp_params* params, int n_devices, const int* devices, pp_group** out) {
  if (!params || !out || n_devices < 1) { pp::set_last_error("pp_group_create: bad arguments"); return PP_ERR_INVALID; }
  *out = nullptr;
  pp_group* g = new pp_group();
  g->n = n_devices;
  for (int d = 0; d < n_devices; ++d) g->devices.push_back(devices ? devices[d] : d);
  if (!g->nccl.load()) {
    pp::set_last_error("pp_group_create: libnccl.so.2 not found or incomplete");
    delete g;
    return PP_ERR_NCCL;
  }
  g->h.assign(n_devices, nullptr);
  for (int d = 0; d < n_devices; ++d) {
    int rc = pp_create(params, g->devices[d], &g->h[d]);
    if (rc != PP_OK) { pp_group_destroy(g); return rc; }
  }
  g->comms.assign(n_devices, nullptr);
  ncclResult_t r = g->nccl.CommInitAll(g->comms.data(), n_devices, g->devices.data());
  if (r != 0) {
    pp::set_last_error(std::string("ncclCommInitAll: ") + g->nccl.GetErrorString(r));
    g->comms.clear();
    pp_group_destroy(g);
    return PP_ERR_NCCL;
  }
  g->d_send.assign(n_devices, nullptr);
  g->d_all.assign(n_devices, nullptr);
  g->send_cap.assign(n_devices, 0);
  g->all_cap.assign(n_devices, 0);
  *out = g;
  return PP_OK;
}

int pp_group_destroy(pp_group* g) {
  if (!g) return PP_OK;
  for (size_t d = 0; d < g->comms.size(); ++d)
    if (g->comms[d]) g->nccl.CommDestroy(g->comms[d]);
  for (size_t d = 0; d < g->d_send.size(); ++d) {
    cudaSetDevice(g->devices[d]);
    cudaFree(g->d_send[d]);
    cudaFree(g->d_all[d]);
  }
  for (pp_handle* h : g->h) pp_destroy(h);
  delete g;
  return PP_OK;
}

int pp_group_size(const pp_group* g, int* out) {
  if (!g || !out) return PP_ERR_INVALID;
  *out = g->n;
  return PP_OK;
}

int pp_group_set_grid(pp_group* g, const int8_t* data_host, int32_t width, int32_t height, double res, int layout) {
  if (!g) return PP_ERR_INVALID;
  for (int d = 0; d < g->n; ++d) {
    int rc = pp_set_grid(g->h[d], data_host, width, height, res, layout);
    if (rc != PP_OK) return rc;
  }
  return PP_OK;
}

int pp_group_collide_batch(pp_group* g, int64_t n, const float* poses_host, uint8_t* out_flags_host) {
  if (!g || n < 0 || (n > 0 && (!poses_host || !out_flags_host))) return PP_ERR_INVALID;
  if (n == 0) return PP_OK;
  const int G = g->n;
  const int64_t per = (n + G - 1) / G;   // equal count per rank for the all-gather; the tail is padding
  int rc = ensure_gather(g, static_cast<size_t>(per));
  if (rc) return rc;
  // 1. upload each shard and launch its kernel (asynchronous per device)
  std::vector<float*> d_poses(G, nullptr);
  for (int d = 0; d < G; ++d) {
    const int64_t lo = std::min<int64_t>(n, d * per), hi = std::min<int64_t>(n, lo + per), m = hi - lo;
    pp_handle* h = g->h[d];
    PP_CUDA(cudaSetDevice(g->devices[d]));
    if (m > h->stage_cap) {
      cudaFree(h->d_poses); cudaFree(h->d_flags);
      h->d_poses = nullptr; h->d_flags = nullptr; h->stage_cap = 0;
      PP_CUDA(cudaMalloc(&h->d_poses, static_cast<size_t>(per) * 3 * sizeof(float)));
      PP_CUDA(cudaMalloc(&h->d_flags, static_cast<size_t>(per)));
      h->stage_cap = per;
    }
    PP_CUDA(cudaMemsetAsync(g->d_send[d], 0, static_cast<size_t>(per), h->stream));
    if (m > 0) {
      PP_CUDA(cudaMemcpyAsync(h->d_poses, poses_host + 3 * lo, static_cast<size_t>(m) * 3 * sizeof(float),
                              cudaMemcpyHostToDevice, h->stream));
      rc = pp_collide_batch_dev(h, m, h->d_poses, g->d_send[d]);
      if (rc) return rc;
    }
  }
  // 2. gather the flags: afterwards every device holds all shards in input order
  rc = all_gather(g, static_cast<size_t>(per));
  if (rc) return rc;
  // 3. the host copy comes from device 0
  PP_CUDA(cudaSetDevice(g->devices[0]));
  for (int d = 0; d < G; ++d) {
    const int64_t lo = std::min<int64_t>(n, d * per), hi = std::min<int64_t>(n, lo + per), m = hi - lo;
    if (m > 0)
      PP_CUDA(cudaMemcpyAsync(out_flags_host + lo, g->d_all[0] + static_cast<size_t>(d) * per, static_cast<size_t>(m),
                              cudaMemcpyDeviceToHost, g->h[0]->stream));
  }
  for (int d = 0; d < G; ++d) {
    PP_CUDA(cudaSetDevice(g->devices[d]));
    PP_CUDA(cudaStreamSynchronize(g->h[d]->stream));
  }
  return PP_OK;
}

int pp_group_plan_batch(pp_group* g, int32_t n, const pp_query* queries, pp_result* results) {
  if (!g || n < 0 || (n > 0 && (!queries || !results))) return PP_ERR_INVALID;
  if (n == 0) return PP_OK;
  const int G = g->n;
  const int per = (n + G - 1) / G;
  bool want_nodes = false;
  int cap_path = 2;
  for (int k = 0; k < n; ++k) {
    want_nodes |= results[k].node_state || results[k].node_parent || results[k].node_closed_seq || results[k].expansion_ids;
    if (results[k].cap_path < 0) { pp::set_last_error("pp_group_plan_batch: negative capacity"); return PP_ERR_INVALID; }
    cap_path = std::max(cap_path, results[k].cap_path);
  }
  cap_path = std::min(cap_path, 1 << 12);
  if (want_nodes) {
    // whole trees are a debugging aid and stay per device: every device plans its contiguous block through the
    // host-pointer call (one host thread each, the call is synchronous) and the results are complete after the join
    std::vector<int> rcs(G, PP_OK);
    std::vector<std::string> errs(G);
    std::vector<std::thread> th;
    for (int d = 0; d < G; ++d) {
      const int lo = std::min(n, d * per), hi = std::min(n, lo + per);
      if (hi <= lo) continue;
      th.emplace_back([=, &rcs, &errs]() {
        rcs[d] = pp_plan_batch(g->h[d], hi - lo, queries + lo, results + lo);
        if (rcs[d] != PP_OK) errs[d] = pp_last_error();
      });
    }
    for (auto& t : th) t.join();
    int ret = PP_OK;
    for (int d = 0; d < G; ++d)
      if (rcs[d] != PP_OK) { pp::set_last_error(errs[d]); if (rcs[d] != PP_ERR_CAPACITY) return rcs[d]; ret = rcs[d]; }
    return ret;
  }
  // 1. every device plans its contiguous block of queries and leaves fixed-stride records (summary + path + profile,
  //    LP:503-583) in ITS send buffer: asynchronous launches, no host thread per device, nothing crosses PCIe
  const size_t stride = pp_plan_record_stride(cap_path);
  const size_t per_bytes = stride * static_cast<size_t>(per);
  int rc = ensure_gather(g, per_bytes);
  if (rc) return rc;
  for (int d = 0; d < G; ++d) {
    const int lo = std::min(n, d * per), hi = std::min(n, lo + per);
    PP_CUDA(cudaSetDevice(g->devices[d]));
    if (hi - lo < per)     // the padding records of the last block(s)
      PP_CUDA(cudaMemsetAsync(g->d_send[d] + stride * (hi - lo), 0, stride * (per - (hi - lo)), g->h[d]->stream));
    if (hi > lo) {
      rc = pp_plan_batch_dev(g->h[d], hi - lo, queries + lo, cap_path, g->d_send[d]);
      if (rc) return rc;
    }
  }
  // 2. one ncclAllGather of the records over NVLink: afterwards every device holds all results in input order
  rc = all_gather(g, per_bytes);
  if (rc) return rc;
  // 3. the host copy comes from device 0, once
  std::vector<unsigned char> all(per_bytes * G);
  PP_CUDA(cudaSetDevice(g->devices[0]));
  PP_CUDA(cudaMemcpyAsync(all.data(), g->d_all[0], per_bytes * G, cudaMemcpyDeviceToHost, g->h[0]->stream));
  for (int d = 0; d < G; ++d) {
    PP_CUDA(cudaSetDevice(g->devices[d]));
    PP_CUDA(cudaStreamSynchronize(g->h[d]->stream));
  }
  return pp_plan_records_unpack(all.data(), n, cap_path, results);
}

// device d's copy of the records gathered by the last pp_group_plan_batch (n x pp_plan_record_stride(cap_path)
// bytes, input order) -- for callers that keep working on the device
int pp_group_records_dev(pp_group* g, int device_index, const void** out_records_dev) {
  if (!g || !out_records_dev || device_index < 0 || device_index >= g->n) return PP_ERR_INVALID;
  *out_records_dev = g->d_all[device_index];
  return PP_OK;
}

}  // extern "C"
