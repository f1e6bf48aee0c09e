// pp_api.cu -- the extern "C" boundary of include/prius_planner.h (host side, C++).
//
// Everything here is argument checking, memory staging and stream / event plumbing; the
// arithmetic lives in the kernels.  There is no CPU fallback: without a CUDA device
// pp_create fails.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "pp_common.cuh"

namespace pp {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }

int philox_kat(pp_handle* h, const uint32_t c[4], const uint32_t k[2], uint32_t out[4]);

static int ceil_tol(double q) { return static_cast<int>(std::ceil(q - 1e-9)); }

// DESIGN.md "ORIENTED footprint lattice": nu x nv points spanning length x width, spacing
// <= one cell.  Same expressions as the oracle's make_lattice.
Lattice make_lattice(double car_length, double car_width, double res) {
  Lattice l;
  const double lu = car_length / res, lv = car_width / res;
  l.nu = ceil_tol(lu) + 1;
  l.nv = ceil_tol(lv) + 1;
  l.su = l.nu > 1 ? lu / (l.nu - 1) : 0.0;
  l.sv = l.nv > 1 ? lv / (l.nv - 1) : 0.0;
  l.hu = (l.nu - 1) * 0.5;
  l.hv = (l.nv - 1) * 0.5;
  return l;
}
bool lattice_supported(const Lattice& l) {
  const double du = l.su * l.hu, dv = l.sv * l.hv;
  return std::sqrt(du * du + dv * dv) + 1.0 < 47.0 && l.nu >= 1 && l.nv >= 1;
}

static int ensure_stage(pp_handle* h, int64_t n) {
  if (n <= h->stage_cap) return PP_OK;
  cudaFree(h->d_poses);
  cudaFree(h->d_flags);
  h->d_poses = nullptr; h->d_flags = nullptr; h->stage_cap = 0;
  PP_CUDA(cudaMalloc(&h->d_poses, static_cast<size_t>(n) * 3 * sizeof(float)));
  PP_CUDA(cudaMalloc(&h->d_flags, static_cast<size_t>(n)));
  h->stage_cap = n;
  return PP_OK;
}

}  // namespace pp

using namespace pp;

extern "C" {

int pp_abi_version(void) { return PP_ABI_VERSION; }
const char* pp_last_error(void) { return g_last_error.c_str(); }

// full_sim.launch:77-79 (75 m / 0.25 m = 300 cells), :90-97; local_planner.hpp:141;
// footprint from prius.urdf through LP:328-334
int pp_default_params(pp_params* p) {
  if (!p) return PP_ERR_INVALID;
  std::memset(p, 0, sizeof(*p));
  p->costmap_res = 0.25;
  p->costmap_width = 300;
  p->costmap_height = 300;
  p->collision_buffer_distance = 1.0;
  p->time_step_ms = 150.0;
  p->velocity_res = 0.1;
  p->heading_res = 0.005;
  p->heading_diff = 0.065;
  p->goal_pos_tolerance = 0.1;
  p->goal_speed_tolerance = 0.5;
  p->search_res = 0.25;
  p->max_velocity = 15.0;
  p->car_length = 4.7;
  p->car_width = 1.586;
  p->collision_mode = PP_COLLISION_AS_SHIPPED;
  p->max_expansions = 100000;
  // the reference's plan-time guard is 1 s of ros::Time (LP:33-39); on one core of a 2018-class desktop its O(N)
  // std::find per candidate reaches ~62 k nodes / ~810 expansions in that second.  60 000 is
  // the deterministic stand-in for it, here and in the ROS shim.
  p->max_nodes = 60000;
  p->rrt_samples_per_round = 256;
  p->rrt_max_samples = 1 << 20;
  p->rrt_goal_bias = 0.05;
  p->rrt_goal_tolerance = 1.0;
  return PP_OK;
}

static int create_resources(pp_handle* h) {
  cudaDeviceProp prop;
  PP_CUDA(cudaGetDeviceProperties(&prop, h->device));
  h->sm_count = prop.multiProcessorCount;
  PP_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  PP_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  PP_CUDA(cudaStreamCreateWithFlags(&h->out_stream, cudaStreamNonBlocking));
  for (auto& ev : h->ev) PP_CUDA(cudaEventCreate(&ev));
  PP_CUDA(cudaMalloc(&h->d_count, 64));
  PP_CUDA(cudaMemset(h->d_count, 0, 64));
  PP_CUDA(cudaMallocHost(&h->h_count, 64));
  std::memset(h->h_count, 0, 64);
  return PP_OK;
}

int pp_create(const pp_params* params, int device, pp_handle** out) {
  if (!params || !out) { set_last_error("pp_create: null argument"); return PP_ERR_INVALID; }
  *out = nullptr;
  PP_REQUIRE(params->costmap_width > 0 && params->costmap_height > 0, "pp_create: empty grid");
  PP_REQUIRE(params->costmap_width <= 32768 && params->costmap_height <= 32768, "pp_create: grid too large");
  PP_REQUIRE(params->costmap_res > 0, "pp_create: costmap_res must be > 0");
  PP_REQUIRE(params->collision_mode >= 0 && params->collision_mode <= 2, "pp_create: bad collision_mode");
  PP_REQUIRE(params->car_length > 0 && params->car_width > 0, "pp_create: bad footprint");
  const Lattice lat = make_lattice(params->car_length, params->car_width, params->costmap_res);
  if (!lattice_supported(lat)) {
    set_last_error("pp_create: footprint spans more than 94 cells; not supported by the tile window");
    return PP_ERR_UNSUPPORTED;
  }
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0) {
    set_last_error(std::string("pp_create: no CUDA device (") + cudaGetErrorString(e) +
                   "); this library has no CPU fallback");
    return PP_ERR_CUDA;
  }
  PP_REQUIRE(device >= 0 && device < n_dev, "pp_create: bad device index");
  PP_CUDA(cudaSetDevice(device));
  pp_handle* h = new pp_handle();
  h->P = *params;
  h->device = device;
  h->lat = lat;
  h->grid.W = params->costmap_width;
  h->grid.H = params->costmap_height;
  h->grid.res = params->costmap_res;
  const int rc = create_resources(h);
  if (rc != PP_OK) {          // streams / events / buffers created so far are released again
    pp_destroy(h);
    return rc;
  }
  *out = h;
  return PP_OK;
}

int pp_destroy(pp_handle* h) {
  if (!h) return PP_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  plan_free(h);
  rrt_free(h);
  grid_free(h);
  cudaFree(h->d_poses);
  cudaFree(h->d_flags);
  cudaFree(h->d_q16);
  cudaFree(h->d_visited);
  cudaFree(h->d_nn_best);
  cudaFree(h->d_count);
  if (h->h_count) cudaFreeHost(h->h_count);
  cudaFree(h->scratch);
  for (auto& ev : h->ev) if (ev) cudaEventDestroy(ev);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->out_stream) cudaStreamDestroy(h->out_stream);
  for (int k = 0; k < h->chunk_ev_n; ++k) cudaEventDestroy(h->chunk_ev[k]);
  delete[] h->chunk_ev;
  delete h;
  return PP_OK;
}

int pp_get_params(const pp_handle* h, pp_params* out) {
  if (!h || !out) return PP_ERR_INVALID;
  *out = h->P;
  return PP_OK;
}

int pp_set_collision_mode(pp_handle* h, int mode) {
  if (!h) return PP_ERR_INVALID;
  PP_REQUIRE(mode >= 0 && mode <= 2, "pp_set_collision_mode: bad mode");
  h->P.collision_mode = mode;
  return PP_OK;
}

// nav_msgs/MapMetaData.resolution is float32: a launch-file 0.1 arrives as 0.100000001490116.  The reference
// never looks at the message's resolution at all (LP:425-435 uses its own parameter), so the check only
// guards against a grid of a different scale: equal as float32, or within 1e-6 relative.
static bool same_resolution(double res, double want) {
  return static_cast<float>(res) == static_cast<float>(want) || std::fabs(res - want) <= 1e-6 * std::fabs(want);
}

static int set_grid_common(pp_handle* h, const int8_t* data, bool on_device, int32_t width, int32_t height,
                           double res, int layout) {
  if (!h || !data) { set_last_error("pp_set_grid: null argument"); return PP_ERR_INVALID; }
  PP_REQUIRE(width == h->P.costmap_width && height == h->P.costmap_height,
             "pp_set_grid: width/height differ from the handle's costmap_width/height");
  PP_REQUIRE(same_resolution(res, h->P.costmap_res), "pp_set_grid: res differs from the handle's costmap_res");
  PP_REQUIRE(layout == PP_GRID_OCCUPANCY_MSG || layout == PP_GRID_CSPACE, "pp_set_grid: bad layout");
  PP_CUDA(cudaSetDevice(h->device));
  return grid_upload(h, data, on_device, layout);
}
int pp_set_grid(pp_handle* h, const int8_t* data_host, int32_t width, int32_t height, double res, int layout) {
  return set_grid_common(h, data_host, false, width, height, res, layout);
}
int pp_set_grid_dev(pp_handle* h, const int8_t* data_dev, int32_t width, int32_t height, double res, int layout) {
  return set_grid_common(h, data_dev, true, width, height, res, layout);
}

int pp_get_cspace(pp_handle* h, int8_t* out_host, size_t capacity) {
  if (!h || !out_host) return PP_ERR_INVALID;
  if (!h->has_grid) { set_last_error("pp_get_cspace: no grid"); return PP_ERR_NO_GRID; }
  const size_t cells = static_cast<size_t>(h->grid.W) * h->grid.H;
  if (capacity < cells) { set_last_error("pp_get_cspace: buffer too small"); return PP_ERR_CAPACITY; }
  PP_CUDA(cudaSetDevice(h->device));
  PP_CUDA(cudaMemcpyAsync(out_host, h->grid.cspace, cells, cudaMemcpyDeviceToHost, h->stream));
  PP_CUDA(cudaStreamSynchronize(h->stream));
  return PP_OK;
}

static int require_grid(pp_handle* h, const char* who) {
  if (!h) return PP_ERR_INVALID;
  if (h->P.collision_mode != PP_COLLISION_AS_SHIPPED && !h->has_grid) {
    set_last_error(std::string(who) + ": pp_set_grid has not been called");
    return PP_ERR_NO_GRID;
  }
  return PP_OK;
}

int pp_collide_batch_dev(pp_handle* h, int64_t n, const float* poses_dev, uint8_t* out_flags_dev) {
  int rc = require_grid(h, "pp_collide_batch_dev");
  if (rc) return rc;
  PP_REQUIRE(n >= 0 && (n == 0 || (poses_dev && out_flags_dev)), "pp_collide_batch_dev: bad arguments");
  PP_CUDA(cudaSetDevice(h->device));
  return collide_batch_dev(h, n, poses_dev, out_flags_dev);
}

int pp_collide_batch_bits_dev(pp_handle* h, int64_t n, const float* poses_dev, uint32_t* out_bits_dev) {
  int rc = require_grid(h, "pp_collide_batch_bits_dev");
  if (rc) return rc;
  PP_REQUIRE(n >= 0 && (n == 0 || (poses_dev && out_bits_dev)), "pp_collide_batch_bits_dev: bad arguments");
  PP_CUDA(cudaSetDevice(h->device));
  if (h->P.collision_mode == PP_COLLISION_ORIENTED)     // the kernel writes the bits itself
    return collide_batch_dev(h, n, poses_dev, reinterpret_cast<uint8_t*>(out_bits_dev), true);
  rc = ensure_stage(h, n);                              // the other modes: byte flags into the staging buffer, then packed
  if (rc) return rc;
  rc = collide_batch_dev(h, n, poses_dev, h->d_flags);
  if (rc) return rc;
  return pack_flags_dev(h, n, h->d_flags, out_bits_dev);
}

// Host-pointer entry points: H2D, kernels and D2H pipelined in chunks on three streams so the PCIe copies overlap
// the kernels.  `poses_host` (float32 triples) or `q16_host` (int16 triples + `quant`, dequantised on the device).
static int collide_batch_host(pp_handle* h, int64_t n, const float* poses_host, const int16_t* q16_host,
                              const pp_pose_quant* quant, uint8_t* out_flags_host) {
  PP_CUDA(cudaSetDevice(h->device));
  int rc = ensure_stage(h, n);
  if (rc) return rc;
  if (q16_host && n > h->q16_cap) {
    cudaFree(h->d_q16);
    h->d_q16 = nullptr; h->q16_cap = 0;
    PP_CUDA(cudaMalloc(&h->d_q16, static_cast<size_t>(n) * 3 * sizeof(int16_t)));
    h->q16_cap = n;
  }
  // Chunk schedule: large copies first (the DMA engine runs closest to the link rate on few large transfers), ever
  // smaller ones towards the end (what follows the last upload -- its kernel and its download -- is not overlapped
  // with anything): each chunk a third of what is left, between 256 K and 4 M poses (halves, 8 M or 16 M at the top and
  // 64 K at the bottom all measured within 0.5 %; fixed 1 M chunks: +1.3 %).  PP_COLLIDE_HOST_CHUNK = fixed size.
  static const int64_t fixed_chunk = getenv("PP_COLLIDE_HOST_CHUNK") ? std::max<int64_t>(1024, atoll(getenv("PP_COLLIDE_HOST_CHUNK"))) : 0;
  std::vector<int64_t> sizes;
  for (int64_t left = n; left > 0;) {
    int64_t m = fixed_chunk ? fixed_chunk : std::min<int64_t>(4 << 20, std::max<int64_t>(256 << 10, left / 3)) & ~int64_t(1023);
    if (m >= left || left - m < (128 << 10)) m = left;
    sizes.push_back(m);
    left -= m;
  }
  const int64_t n_chunks = static_cast<int64_t>(sizes.size());
  cudaStream_t in = h->copy_stream, ks = h->stream, out = h->out_stream;
  if (2 * n_chunks > h->chunk_ev_n) {
    for (int k = 0; k < h->chunk_ev_n; ++k) cudaEventDestroy(h->chunk_ev[k]);
    delete[] h->chunk_ev;
    h->chunk_ev_n = static_cast<int>(2 * n_chunks);
    h->chunk_ev = new cudaEvent_t[h->chunk_ev_n];
    for (int k = 0; k < h->chunk_ev_n; ++k) PP_CUDA(cudaEventCreateWithFlags(&h->chunk_ev[k], cudaEventDisableTiming));
  }
  // three streams, one per engine: uploads never wait for kernels or downloads, the kernel of
  // chunk c waits only for its own upload, the download of chunk c only for its own kernel
  PP_CUDA(cudaEventRecord(h->ev[0], in));
  rc = collide_counter_reset(h);
  if (rc) return rc;
  PP_CUDA(cudaEventRecord(h->ev[2], ks));
  int64_t lo = 0;
  for (int64_t c = 0; c < n_chunks; lo += sizes[c], ++c) {
    const int64_t m = sizes[c];
    cudaEvent_t up = h->chunk_ev[2 * c], done = h->chunk_ev[2 * c + 1];
    if (q16_host)
      PP_CUDA(cudaMemcpyAsync(h->d_q16 + 3 * lo, q16_host + 3 * lo, static_cast<size_t>(m) * 3 * sizeof(int16_t),
                              cudaMemcpyHostToDevice, in));
    else
      PP_CUDA(cudaMemcpyAsync(h->d_poses + 3 * lo, poses_host + 3 * lo, static_cast<size_t>(m) * 3 * sizeof(float),
                              cudaMemcpyHostToDevice, in));
    PP_CUDA(cudaEventRecord(up, in));
    PP_CUDA(cudaStreamWaitEvent(ks, up, 0));
    if (q16_host) {
      rc = dequant_poses_dev(h, m, h->d_q16 + 3 * lo, *quant, h->d_poses + 3 * lo);
      if (rc) return rc;
    }
    rc = collide_launch(h, m, h->d_poses + 3 * lo, h->d_flags + lo);
    if (rc) return rc;
    PP_CUDA(cudaEventRecord(done, ks));
    PP_CUDA(cudaStreamWaitEvent(out, done, 0));
    PP_CUDA(cudaMemcpyAsync(out_flags_host + lo, h->d_flags + lo, static_cast<size_t>(m), cudaMemcpyDeviceToHost, out));
  }
  PP_CUDA(cudaEventRecord(h->ev[3], ks));
  rc = collide_counter_fetch(h);
  if (rc) return rc;
  PP_CUDA(cudaEventRecord(h->ev[1], out));
  PP_CUDA(cudaStreamSynchronize(in));
  PP_CUDA(cudaStreamSynchronize(ks));
  PP_CUDA(cudaStreamSynchronize(out));
  if (h->P.collision_mode != PP_COLLISION_ORIENTED) h->last_narrow = n;
  float ms = 0;
  if (cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]) == cudaSuccess) h->phase_ms[PP_PHASE_H2D] = ms;
  else cudaGetLastError();
  return PP_OK;
}

int pp_collide_batch(pp_handle* h, int64_t n, const float* poses_host, uint8_t* out_flags_host) {
  int rc = require_grid(h, "pp_collide_batch");
  if (rc) return rc;
  PP_REQUIRE(n >= 0 && (n == 0 || (poses_host && out_flags_host)), "pp_collide_batch: bad arguments");
  if (n == 0) return PP_OK;
  return collide_batch_host(h, n, poses_host, nullptr, nullptr, out_flags_host);
}

int pp_collide_batch_q16(pp_handle* h, int64_t n, const int16_t* poses_q16_host, const pp_pose_quant* quant,
                         uint8_t* out_flags_host) {
  int rc = require_grid(h, "pp_collide_batch_q16");
  if (rc) return rc;
  PP_REQUIRE(n >= 0 && quant && (n == 0 || (poses_q16_host && out_flags_host)), "pp_collide_batch_q16: bad arguments");
  if (n == 0) return PP_OK;
  return collide_batch_host(h, n, nullptr, poses_q16_host, quant, out_flags_host);
}

int pp_pack_flags_dev(pp_handle* h, int64_t n, const uint8_t* flags_dev, uint32_t* out_bits_dev) {
  if (!h) return PP_ERR_INVALID;
  PP_REQUIRE(n >= 0 && (n == 0 || (flags_dev && out_bits_dev)), "pp_pack_flags_dev: bad arguments");
  PP_CUDA(cudaSetDevice(h->device));
  return pack_flags_dev(h, n, flags_dev, out_bits_dev);
}
int pp_unpack_flags_dev(pp_handle* h, int64_t n, const uint32_t* bits_dev, uint8_t* out_flags_dev) {
  if (!h) return PP_ERR_INVALID;
  PP_REQUIRE(n >= 0 && (n == 0 || (bits_dev && out_flags_dev)), "pp_unpack_flags_dev: bad arguments");
  PP_CUDA(cudaSetDevice(h->device));
  return unpack_flags_dev(h, n, bits_dev, out_flags_dev);
}

int pp_collide_edges_dev(pp_handle* h, int64_t n, const float* edges_dev, uint8_t* out_flags_dev) {
  int rc = require_grid(h, "pp_collide_edges_dev");
  if (rc) return rc;
  PP_REQUIRE(n >= 0 && (n == 0 || (edges_dev && out_flags_dev)), "pp_collide_edges_dev: bad arguments");
  PP_CUDA(cudaSetDevice(h->device));
  return collide_edges_dev(h, n, edges_dev, out_flags_dev);
}

int pp_footprint_points(pp_handle* h, double x, double y, double yaw, double* out_xy_host, int32_t capacity,
                        int32_t* n_points) {
  if (!h || !out_xy_host || !n_points || capacity < 0) return PP_ERR_INVALID;
  PP_CUDA(cudaSetDevice(h->device));
  // one allocation: [count | pad | 2*capacity doubles]; debug entry point, not on the hot path
  char* blk = nullptr;
  const size_t bytes = 16 + sizeof(double) * 2 * static_cast<size_t>(capacity);
  PP_CUDA(cudaMalloc(&blk, bytes));
  int* dn = reinterpret_cast<int*>(blk);
  double* d = reinterpret_cast<double*>(blk + 16);
  int rc = footprint_points_dev(h, x, y, yaw, d, capacity, dn);
  int n = 0;
  cudaError_t e = cudaSuccess;
  if (rc == PP_OK) e = cudaMemcpyAsync(&n, dn, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
  if (rc == PP_OK && e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (rc == PP_OK && e == cudaSuccess && n > 0) {
    e = cudaMemcpyAsync(out_xy_host, d, sizeof(double) * 2 * static_cast<size_t>(n), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  }
  cudaFree(blk);
  if (rc != PP_OK) return rc;
  PP_CUDA(e);
  *n_points = n;
  return PP_OK;
}

int pp_sample_poses_dev(pp_handle* h, uint64_t seed, uint32_t stream, uint64_t first_index, int64_t n, float x_lo,
                        float x_hi, float y_lo, float y_hi, float* out_poses_dev) {
  if (!h || n < 0 || (n > 0 && !out_poses_dev)) return PP_ERR_INVALID;
  PP_CUDA(cudaSetDevice(h->device));
  return sample_poses_dev(h, seed, stream, first_index, n, x_lo, x_hi, y_lo, y_hi, out_poses_dev);
}

int pp_philox4x32_10(pp_handle* h, const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]) {
  if (!h || !counter || !key || !out) return PP_ERR_INVALID;
  PP_CUDA(cudaSetDevice(h->device));
  return philox_kat(h, counter, key, out);
}

int pp_nearest_dev(pp_handle* h, int64_t n_nodes, const float* node_xy_dev, int64_t n_query,
                   const float* query_xy_dev, int32_t* out_index_dev, float* out_dist2_dev) {
  if (!h || n_nodes < 0 || n_query < 0) return PP_ERR_INVALID;
  if (n_query > 0 && (!query_xy_dev || !out_index_dev || (n_nodes > 0 && !node_xy_dev))) return PP_ERR_INVALID;
  PP_CUDA(cudaSetDevice(h->device));
  return nearest_dev(h, n_nodes, node_xy_dev, n_query, query_xy_dev, out_index_dev, out_dist2_dev);
}

int pp_find_exact_dev(pp_handle* h, int64_t n_nodes, const double* node_keys_dev, int64_t n_query,
                      const double* query_keys_dev, int32_t* out_index_dev) {
  if (!h || n_nodes < 0 || n_query < 0) return PP_ERR_INVALID;
  if (n_query > 0 && (!query_keys_dev || !out_index_dev || (n_nodes > 0 && !node_keys_dev))) return PP_ERR_INVALID;
  PP_CUDA(cudaSetDevice(h->device));
  return find_exact_dev(h, n_nodes, node_keys_dev, n_query, query_keys_dev, out_index_dev);
}

int pp_plan(pp_handle* h, const pp_query* query, pp_result* result) { return pp_plan_batch(h, 1, query, result); }

int pp_set_grid_bank_dev(pp_handle* h, int32_t n, const int8_t* grids_dev, int32_t width, int32_t height, double res,
                         int layout) {
  if (!h) return PP_ERR_INVALID;
  PP_REQUIRE(n >= 1 && n <= 65535 && grids_dev, "pp_set_grid_bank_dev: 1 <= n <= 65535 grids on the device");
  PP_REQUIRE(width == h->P.costmap_width && height == h->P.costmap_height && same_resolution(res, h->P.costmap_res),
             "pp_set_grid_bank_dev: grid shape differs from the handle's parameters");
  PP_REQUIRE(layout == PP_GRID_OCCUPANCY_MSG || layout == PP_GRID_CSPACE, "pp_set_grid_bank_dev: unknown layout");
  return grid_bank_upload(h, n, grids_dev, layout);
}

int pp_plan_batch_banked(pp_handle* h, int32_t n, const pp_query* queries, pp_result* results) {
  if (!h) return PP_ERR_INVALID;
  PP_REQUIRE(n >= 0 && (n == 0 || (queries && results)), "pp_plan_batch_banked: bad arguments");
  if (n == 0) return PP_OK;
  if (h->bank.n < 1) {
    set_last_error("pp_plan_batch_banked: pp_set_grid_bank_dev has not been called");
    return PP_ERR_NO_GRID;
  }
  for (int32_t k = 0; k < n; ++k)
    PP_REQUIRE(queries[k].reserved < static_cast<uint32_t>(h->bank.n), "pp_plan_batch_banked: query.reserved (grid index) outside the bank");
  PP_CUDA(cudaSetDevice(h->device));
  return plan_batch(h, n, queries, results, true);
}

int pp_plan_batch(pp_handle* h, int32_t n, const pp_query* queries, pp_result* results) {
  int rc = require_grid(h, "pp_plan_batch");
  if (rc) return rc;
  PP_REQUIRE(n >= 0 && (n == 0 || (queries && results)), "pp_plan_batch: bad arguments");
  if (n == 0) return PP_OK;
  PP_CUDA(cudaSetDevice(h->device));
  return plan_batch(h, n, queries, results);
}

size_t pp_plan_record_stride(int32_t cap_path) {
  if (cap_path < 0) return 0;
  return (sizeof(pp_plan_record_header) + static_cast<size_t>(cap_path) * 44 + 15) & ~size_t(15);
}

int pp_plan_batch_dev(pp_handle* h, int32_t n, const pp_query* queries_host, int32_t cap_path, void* out_records_dev) {
  int rc = require_grid(h, "pp_plan_batch_dev");
  if (rc) return rc;
  PP_REQUIRE(n >= 0 && (n == 0 || (queries_host && out_records_dev)), "pp_plan_batch_dev: bad arguments");
  PP_REQUIRE((reinterpret_cast<uintptr_t>(out_records_dev) & 15) == 0, "pp_plan_batch_dev: records must be 16-byte aligned");
  if (n == 0) return PP_OK;
  PP_CUDA(cudaSetDevice(h->device));
  return plan_batch(h, n, queries_host, nullptr, false, static_cast<unsigned char*>(out_records_dev), cap_path);
}

// host helper: records (host memory, e.g. a copy of the gathered device block) -> the caller's pp_result arrays
int pp_plan_records_unpack(const void* records_host, int32_t n, int32_t cap_path, pp_result* results) {
  PP_REQUIRE(n >= 0 && cap_path >= 0 && (n == 0 || (records_host && results)), "pp_plan_records_unpack: bad arguments");
  const size_t stride = pp_plan_record_stride(cap_path);
  int ret = PP_OK;
  for (int32_t k = 0; k < n; ++k) {
    const unsigned char* rec = static_cast<const unsigned char*>(records_host) + stride * k;
    pp_plan_record_header hd;
    std::memcpy(&hd, rec, sizeof(hd));
    pp_result& r = results[k];
    PP_REQUIRE(r.cap_path >= 0, "pp_plan_records_unpack: negative capacity in a pp_result");
    r.status = hd.status; r.n_expansions = hd.n_expansions; r.n_nodes = hd.n_nodes; r.n_open = hd.n_open;
    r.n_closed = hd.n_closed; r.n_path = hd.n_path; r.n_profile = hd.n_profile; r.n_candidates = hd.n_candidates;
    r.n_collisions = hd.n_collisions; r.tree_digest = hd.tree_digest;
    const double* path = reinterpret_cast<const double*>(rec + sizeof(hd));
    const double* prof = path + 2 * static_cast<size_t>(cap_path);
    const int32_t* ids = reinterpret_cast<const int32_t*>(prof + 3 * static_cast<size_t>(cap_path));
    const int np = std::min(hd.n_path, std::min(r.cap_path, cap_path));
    const int nm = std::min(hd.n_profile, std::min(r.cap_path, cap_path));
    if (hd.n_path > std::min(r.cap_path, cap_path) && (r.path_xy || r.path_node_ids)) ret = PP_ERR_CAPACITY;
    if (r.path_xy) std::memcpy(r.path_xy, path, sizeof(double) * 2 * np);
    if (r.path_node_ids) std::memcpy(r.path_node_ids, ids, sizeof(int32_t) * np);
    if (r.yaw_rates) std::memcpy(r.yaw_rates, prof, sizeof(double) * nm);
    if (r.durations) std::memcpy(r.durations, prof + cap_path, sizeof(double) * nm);
    if (r.speeds) std::memcpy(r.speeds, prof + 2 * static_cast<size_t>(cap_path), sizeof(double) * nm);
  }
  if (ret == PP_ERR_CAPACITY) set_last_error("pp_plan_records_unpack: a caller-provided output array (or cap_path) is too small");
  return ret;
}

int pp_get_visited(pp_handle* h, uint8_t* out_host, size_t capacity) {
  if (!h || !out_host) return PP_ERR_INVALID;
  const size_t cells = static_cast<size_t>(h->P.costmap_width) * h->P.costmap_height;
  if (capacity < cells) { set_last_error("pp_get_visited: buffer too small"); return PP_ERR_CAPACITY; }
  if (!h->d_visited) { set_last_error("pp_get_visited: no single-query plan has run yet"); return PP_ERR_INVALID; }
  PP_CUDA(cudaSetDevice(h->device));
  PP_CUDA(cudaMemcpyAsync(out_host, h->d_visited, cells, cudaMemcpyDeviceToHost, h->stream));
  PP_CUDA(cudaStreamSynchronize(h->stream));
  return PP_OK;
}

int pp_rrt_plan(pp_handle* h, const pp_query* query, pp_result* result) {
  return pp_rrt_plan_batch(h, 1, query, result);
}
int pp_rrt_plan_batch(pp_handle* h, int32_t n, const pp_query* queries, pp_result* results) {
  int rc = require_grid(h, "pp_rrt_plan_batch");
  if (rc) return rc;
  PP_REQUIRE(n >= 0 && (n == 0 || (queries && results)), "pp_rrt_plan_batch: bad arguments");
  if (n == 0) return PP_OK;
  PP_CUDA(cudaSetDevice(h->device));
  return rrt_plan_batch(h, n, queries, results);
}

int pp_sync(pp_handle* h) {
  if (!h) return PP_ERR_INVALID;
  PP_CUDA(cudaSetDevice(h->device));
  PP_CUDA(cudaStreamSynchronize(h->stream));
  PP_CUDA(cudaStreamSynchronize(h->copy_stream));
  PP_CUDA(cudaStreamSynchronize(h->out_stream));
  return PP_OK;
}
int pp_stream(pp_handle* h, void** out_stream) {
  if (!h || !out_stream) return PP_ERR_INVALID;
  *out_stream = static_cast<void*>(h->stream);
  return PP_OK;
}
int pp_launch_count(const pp_handle* h, uint64_t* out) {
  if (!h || !out) return PP_ERR_INVALID;
  *out = h->launches;
  return PP_OK;
}
int pp_last_phase_ms(pp_handle* h, float* out_ms, int32_t n) {
  if (!h || !out_ms || n < 0) return PP_ERR_INVALID;
  PP_CUDA(cudaSetDevice(h->device));
  PP_CUDA(cudaStreamSynchronize(h->stream));
  float ms = 0;
  if (cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]) == cudaSuccess) h->phase_ms[PP_PHASE_NARROW] = ms;
  else cudaGetLastError();
  if (h->plan_timing_pending) {       // an asynchronous pp_plan_batch_dev: its events have completed by now
    if (cudaEventElapsedTime(&ms, h->ev[8], h->ev[9]) == cudaSuccess) h->phase_ms[PP_PHASE_PLAN] = ms;
    else cudaGetLastError();
    h->plan_timing_pending = false;
  }
  for (int k = 0; k < n && k < PP_PHASE_COUNT; ++k) out_ms[k] = h->phase_ms[k];
  return PP_OK;
}
int pp_last_narrow_count(pp_handle* h, int64_t* out) {
  if (!h || !out) return PP_ERR_INVALID;
  PP_CUDA(cudaSetDevice(h->device));
  PP_CUDA(cudaStreamSynchronize(h->stream));
  if (h->last_narrow < 0) {
    const volatile unsigned long long* pub = reinterpret_cast<const volatile unsigned long long*>(h->h_count);
    h->last_narrow = static_cast<int64_t>(pub[1] - pub[0]);   // {total before the call, total now}
  }
  *out = h->last_narrow;
  return PP_OK;
}
int pp_bench_l2_gather(pp_handle* h, size_t buffer_bytes, double* out_gbps) {
  if (!h || !out_gbps || buffer_bytes < 64) return PP_ERR_INVALID;
  PP_CUDA(cudaSetDevice(h->device));
  return l2_gather_bench(h, buffer_bytes, out_gbps);
}
int pp_set_broad_phase(pp_handle* h, int enabled) {
  if (!h) return PP_ERR_INVALID;
  h->broad_phase = enabled ? 1 : 0;
  return PP_OK;
}

}  // extern "C"
