// nearl_b200 -- C ABI (include/nearl_b200.h) over the sm_100a kernels.
//
// Replaces the pybind11 module /root/reference/src/actions_py.cpp and the CUDA host wrappers
// voxelize_host / trajectory_voxelization_host (/root/reference/src/voxelize.cu:292-415),
// marching_observer_host / observe_frame_host (/root/reference/src/marching_observers.cu:435-553),
// aggregate_host / sum_reduction_host (/root/reference/src/gpuutils.cu:80-134).  Where those
// allocate, copy, synchronise and free on every call, this layer keeps a per-device context
// with grow-only workspaces, one stream, and checks every CUDA status.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <numeric>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/nearl_b200.h"
#include "agg_kernels.cuh"
#include "cell_kernels.cuh"
#include "common.cuh"
#include "pack_kernels.cuh"
#include "superpose_kernels.cuh"
#include "tile_kernels.cuh"

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};
std::atomic<int> g_force_scan{0};    // nrl_set_option("force_scan", 1): never use the cell-list kernels
std::atomic<int> g_agg_pipe_w{0};    // nrl_set_option("agg_pipe_w" / "agg_pipe_stages"): tile width / ring depth of the pipelined K3
std::atomic<int> g_agg_pipe_stages{0};
std::atomic<int> g_agg_entropy_legacy{0};   // nrl_set_option("agg_entropy_legacy"): former histogram form in the staged / pipelined K3 (tests)  // (0 = chosen from the series length; -1 stages = staged one-shot kernel instead)
std::atomic<int> g_crop_stream_min{2048};   // nrl_set_option("crop_stream_min", n): shortest slice that takes the streaming crop (tests: 1)
std::atomic<int> g_crop_legacy{0};   // nrl_set_option("crop_legacy", 1): never use the streaming crop kernel (tests / A-B timing)
std::atomic<int> g_no_fork{0};       // nrl_set_option("no_fork", 1): the launches of a batch run one after the other on its stream (A-B timing)
std::atomic<int> g_no_fuse{0};       // nrl_set_option("no_fuse", 1): one sweep per observable (tests / A-B timing)

int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char *what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return NRL_ERR_CUDA + (int)e;
}
#define NRL_CUDA(call)                                    \
  do {                                                    \
    cudaError_t e__ = (call);                             \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)
#define NRL_TRY(expr)              \
  do {                             \
    int rc__ = (expr);             \
    if (rc__ != NRL_OK) return rc__; \
  } while (0)

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return NRL_OK;
    if (p) { NRL_CUDA(cudaFree(p)); p = nullptr; cap = 0; }
    size_t want = std::max(bytes, (size_t)4096);
    want = (want + (want >> 2) + 255) & ~(size_t)255;   // 25% slack: grow-only arena
    NRL_CUDA(cudaMalloc(&p, want));
    cap = want;
    return NRL_OK;
  }
  template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct Ctx {
  int dev = -1;
  int sms = 148;
  size_t smem_optin = 227 * 1024;
  cudaStream_t stream = nullptr;
  std::mutex mu;
  DevBuf coords, weights, perm, atoms, inv, off, cnt, out, tasks, flist, cand_scr, ser_scr, misc, series;
  DevBuf xyz, wtabs, sels, resid, orders, resflags;
  DevBuf crops;                                  // descriptor blob of a batch: crops | frame offsets | frame -> crop | frame -> source frame
  DevBuf pbase;                                  // per crop frame: payload base
  DevBuf sup_xform;                              // superposition: per frame rotation (9) + centroid (3), double
  DevBuf sup_idx, sup_ref, sup_rmsd;             // superposition: fitted atom indices, reference set, per-frame RMSD
  DevBuf runcnt;                                 // streaming crop: kept atoms per (frame, run)
  DevBuf batoms, bcnt, cstart, grids;            // cell lists: binned frames, cell prefixes, per-frame cell grids | frame list
  bool attr_set[64] = {false};          // "dynamic shared memory opt-in done" per kernel family, indexed by kAttr*
  bool attr_fused[16] = {false};        // the same for the fused observer sweeps, indexed by their flag set
  cudaStream_t user_stream = nullptr;   // nrl_set_stream: run on the caller's stream instead of our own
  bool use_user_stream = false;
  // side streams: the per-mode tile launches of one batch are independent of each other (and the observers do
  // not wait for the voxeliser's normalisation), so they are forked off the work stream and joined at the end
  static constexpr int kAux = 3;
  cudaStream_t aux[kAux] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[kAux] = {nullptr, nullptr, nullptr};
  // optional per-kernel timing (nrl_profile): CUDA events around every launch, resolved at the next sync
  bool prof_on = false;
  struct ProfRec { int cat; cudaEvent_t a, b; };
  std::vector<ProfRec> prof_pending;
  std::vector<cudaEvent_t> prof_pool;
  double prof_ms[NRL_PROFILE_CATEGORIES] = {0};
  long long prof_cnt[NRL_PROFILE_CATEGORIES] = {0};
  cudaStream_t work_stream() const { return use_user_stream ? user_stream : stream; }
  // the workspace above is one set of buffers: a call that runs on another stream than the previous one waits (on
  // the device) for that call's work before it overwrites them -- see WorkspaceTurn
  cudaStream_t ws_last = nullptr;
  cudaEvent_t ws_event = nullptr, join_event = nullptr;
  bool ws_used = false, ws_recorded = false;
};

// Orders successive users of a lane's workspace that run on DIFFERENT streams (the stream-ordered `_dev` twins run on the
// caller's stream and return without a sync; the next call may arrive on another stream, or on the lane's own).  Same
// stream as the previous call: nothing to do, the stream orders them.  Held under Ctx::mu for the length of the call.
struct WorkspaceTurn {
  Ctx &c; cudaStream_t st;
  WorkspaceTurn(Ctx &c_, cudaStream_t st_) : c(c_), st(st_) {
    if (c.ws_used && c.ws_last != st) {
      if (!c.ws_event) cudaEventCreateWithFlags(&c.ws_event, cudaEventDisableTiming);
      if (!c.ws_recorded) cudaEventRecord(c.ws_event, c.ws_last);      // the lane's own stream: alive as long as the lane
      cudaStreamWaitEvent(st, c.ws_event, 0);
    }
  }
  ~WorkspaceTurn() {
    c.ws_used = true;
    c.ws_last = st;
    c.ws_recorded = false;
    if (st != c.stream) {        // a caller's stream may be gone by the next call: take the marker now
      if (!c.ws_event) cudaEventCreateWithFlags(&c.ws_event, cudaEventDisableTiming);
      c.ws_recorded = cudaEventRecord(c.ws_event, st) == cudaSuccess;
    }
  }
};

// slots of Ctx::attr_set: 0-9 scan tile kernels by launch slot, kAttrCell + the same for the cell-list kernels
enum { kAttrAggStaged = 20, kAttrCrop = 21, kAttrMultiTable = 22, kAttrAggPipe = 23, kAttrCell = 32, kAttrBin = 63 };
enum { CAT_PACK_LEGACY = 0, CAT_CROP_PACK = 1, CAT_ATOM_NORM = 2, CAT_TILE0 = 3, CAT_AGGREGATE = 13, CAT_SUMMATION = 14, CAT_MISC = 15, CAT_CELL_BIN = 16, CAT_SUPERPOSE = 17, CAT_TILE_FUSED = 18 };
const char *kCatNames[NRL_PROFILE_CATEGORIES] = {
    "pack_legacy", "crop_pack", "atom_norm", "tile_voxelize", "tile_existence", "tile_direct_count", "tile_distinct_count",
    "tile_mean_distance", "tile_cumulative_weight", "tile_density", "tile_dispersion", "tile_eccentricity",
    "tile_radius_of_gyration", "aggregate", "summation", "misc", "cell_bin", "superpose", "tile_fused"};

struct ScopedProf {
  Ctx &c; cudaStream_t st; int cat; cudaEvent_t a = nullptr, b = nullptr;
  static cudaEvent_t get(Ctx &c) {
    if (!c.prof_pool.empty()) { cudaEvent_t e = c.prof_pool.back(); c.prof_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
  ScopedProf(Ctx &c_, cudaStream_t st_, int cat_) : c(c_), st(st_), cat(cat_) {
    if (c.prof_on) { a = get(c); b = get(c); cudaEventRecord(a, st); }
  }
  ~ScopedProf() {
    if (a) { cudaEventRecord(b, st); c.prof_pending.push_back({cat, a, b}); }
  }
};
void prof_collect(Ctx &c) {
  for (auto &r : c.prof_pending) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      c.prof_ms[r.cat] += ms;
      c.prof_cnt[r.cat] += 1;
    }
    c.prof_pool.push_back(r.a);
    c.prof_pool.push_back(r.b);
  }
  c.prof_pending.clear();
  cudaGetLastError();
}

std::mutex g_ctx_mu;
Ctx *g_ctx[64][NRL_MAX_LANES] = {{nullptr}};

// A lane is an independent (stream, workspace) pair on a device: work submitted on different lanes
// may overlap (upload of one trajectory under the kernels of another).  Lane 0 serves the twins of
// the reference FFI.
int get_ctx(Ctx **out, int lane = 0) {
  if (lane < 0 || lane >= NRL_MAX_LANES) return fail(NRL_ERR_ARGUMENT, "lane out of range");
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(NRL_ERR_NO_DEVICE, std::string("nearl_b200: no usable CUDA device (") + cudaGetErrorString(e) +
                                       "); this library has no CPU fallback");
  }
  if (dev < 0 || dev >= 64) return fail(NRL_ERR_ARGUMENT, "device index out of range");
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  if (!g_ctx[dev][lane]) {
    cudaDeviceProp prop;
    NRL_CUDA(cudaGetDeviceProperties(&prop, dev));
    Ctx *c = new Ctx();
    c->dev = dev;
    c->sms = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    NRL_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    g_ctx[dev][lane] = c;
  }
  *out = g_ctx[dev][lane];
  return NRL_OK;
}

bool valid_obs(int t) {
  static const int ok[] = {1, 2, 3, 11, 12, 13, 14, 15, 16};
  for (int v : ok) if (v == t) return true;
  return false;
}
bool valid_agg(int t) { return t >= 1 && t <= 8; }

int check_dims(const int *dims, long long *G) {
  if (!dims) return fail(NRL_ERR_ARGUMENT, "dims is NULL");
  if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0) return fail(NRL_ERR_SHAPE, "grid dimensions must be positive");
  const long long g = (long long)dims[0] * dims[1] * dims[2];
  if (g >= (1ll << 31)) return fail(NRL_ERR_SHAPE, "grid has 2^31 voxels or more");
  *G = g;
  return NRL_OK;
}

// ---- tile launches ------------------------------------------------------------------------
constexpr int kCandCap = 512;         // candidates held in shared memory (8 KB; 24 KB for the voxeliser's records)
constexpr int kSerCapFloats = 8192;   // series floats held in shared memory (32 KB: 16 frames x 512 voxels)
constexpr size_t kScratchBudget = (size_t)1 << 30;

template <int MODE, int NWT>
int launch_tiles_mode(Ctx &c, const NrlTileParams &p, int grid, size_t smem, cudaStream_t st, int slot) {
  const int attr = NWT > 1 ? kAttrMultiTable : slot;
  if (!c.attr_set[attr]) {
    NRL_CUDA(cudaFuncSetAttribute(nrl_tile_kernel<MODE, NWT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(c.smem_optin - 2048)));
    c.attr_set[attr] = true;
  }
  {
    ScopedProf sp(c, st, CAT_TILE0 + slot);
    nrl_tile_kernel<MODE, NWT><<<grid, NRL_NT, smem, st>>>(p);
  }
  g_launches.fetch_add(1);
  NRL_CUDA(cudaGetLastError());
  return NRL_OK;
}

// mode_key: observable id, NRL_MODE_VOXELIZE, or kVoxelizeMulti (voxeliser with several weight tables)
constexpr int kVoxelizeMulti = -1;
int launch_tiles(Ctx &c, int mode_key, const NrlTileParams &p, int grid, size_t smem, cudaStream_t st) {
  switch (mode_key) {
    case kVoxelizeMulti: return launch_tiles_mode<NRL_MODE_VOXELIZE, NRL_MAX_CH>(c, p, grid, smem, st, 0);
    case NRL_MODE_VOXELIZE: return launch_tiles_mode<NRL_MODE_VOXELIZE, 1>(c, p, grid, smem, st, 0);
    case NRL_OBS_EXISTENCE: return launch_tiles_mode<NRL_OBS_EXISTENCE, 1>(c, p, grid, smem, st, 1);
    case NRL_OBS_DIRECT_COUNT: return launch_tiles_mode<NRL_OBS_DIRECT_COUNT, 1>(c, p, grid, smem, st, 2);
    case NRL_OBS_DISTINCT_COUNT: return launch_tiles_mode<NRL_OBS_DISTINCT_COUNT, 1>(c, p, grid, smem, st, 3);
    case NRL_OBS_MEAN_DISTANCE: return launch_tiles_mode<NRL_OBS_MEAN_DISTANCE, 1>(c, p, grid, smem, st, 4);
    case NRL_OBS_CUMULATIVE_WEIGHT: return launch_tiles_mode<NRL_OBS_CUMULATIVE_WEIGHT, 1>(c, p, grid, smem, st, 5);
    case NRL_OBS_DENSITY: return launch_tiles_mode<NRL_OBS_DENSITY, 1>(c, p, grid, smem, st, 6);
    case NRL_OBS_DISPERSION: return launch_tiles_mode<NRL_OBS_DISPERSION, 1>(c, p, grid, smem, st, 7);
    case NRL_OBS_ECCENTRICITY: return launch_tiles_mode<NRL_OBS_ECCENTRICITY, 1>(c, p, grid, smem, st, 8);
    case NRL_OBS_RADIUS_OF_GYRATION: return launch_tiles_mode<NRL_OBS_RADIUS_OF_GYRATION, 1>(c, p, grid, smem, st, 9);
    default: return fail(NRL_ERR_OBSERVABLE, "The observable type is not supported");
  }
}

// ---- cell-list launches (cell_kernels.cuh) ----------------------------------------------------
struct CellSetup {
  bool on = false;
  NrlFrames frames{};                 // binned frames
  const NrlCellGrid *grids = nullptr; // per frame descriptor
  const int *cell_start = nullptr;
};

bool cell_mode(int mode) {
  return mode == NRL_MODE_VOXELIZE || mode == NRL_OBS_EXISTENCE || mode == NRL_OBS_DIRECT_COUNT ||
         mode == NRL_OBS_MEAN_DISTANCE || mode == NRL_OBS_CUMULATIVE_WEIGHT || mode == NRL_OBS_DENSITY ||
         mode == NRL_OBS_ECCENTRICITY;      // + distinct_count when its values have <= 32 ranks (decided by the caller)
}

// The information entropy is discontinuous in its samples (a histogram): where a sample sits on a bin edge -- always the
// case for sums of discrete per-atom weights, {A - m, A, A, A + m} puts A exactly on one -- the last bit decides, and the last
// bit of a float sum depends on the order of the additions.  Entropy tasks over float observables therefore take the scan
// kernel, which adds the atoms in index order like the reference and is bit-identical to it, whatever the grid.
bool index_order_matters(int mode, int agg) {
  return agg == 7 && (mode == NRL_OBS_MEAN_DISTANCE || mode == NRL_OBS_CUMULATIVE_WEIGHT || mode == NRL_OBS_DENSITY ||
                      mode == NRL_OBS_ECCENTRICITY);
}

template <int MODE, int NWT>
int launch_cell_tiles_mode(Ctx &c, const NrlCellTileParams &p, int grid, size_t smem, cudaStream_t st, int slot) {
  const int attr = kAttrCell + (NWT > 1 ? kAttrMultiTable : slot);
  if (!c.attr_set[attr]) {
    NRL_CUDA(cudaFuncSetAttribute(nrl_cell_tile_kernel<MODE, NWT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(c.smem_optin - 2048)));
    c.attr_set[attr] = true;
  }
  {
    ScopedProf sp(c, st, CAT_TILE0 + slot);
    nrl_cell_tile_kernel<MODE, NWT><<<grid, NRL_NT, smem, st>>>(p);
  }
  g_launches.fetch_add(1);
  NRL_CUDA(cudaGetLastError());
  return NRL_OK;
}

template <int FLAGS>
int launch_cell_fused(Ctx &c, const NrlCellTileParams &p, int grid, size_t smem, cudaStream_t st) {
  static_assert(FLAGS >= 0 && FLAGS < 16, "fused flag set");
  if (!c.attr_fused[FLAGS]) {        // per lane, touched under Ctx::mu like attr_set
    NRL_CUDA(cudaFuncSetAttribute(nrl_cell_tile_kernel<NRL_OBS_FUSED, FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(c.smem_optin - 2048)));
    c.attr_fused[FLAGS] = true;
  }
  {
    ScopedProf sp(c, st, CAT_TILE_FUSED);
    nrl_cell_tile_kernel<NRL_OBS_FUSED, FLAGS><<<grid, NRL_NT, smem, st>>>(p);
  }
  g_launches.fetch_add(1);
  NRL_CUDA(cudaGetLastError());
  return NRL_OK;
}

constexpr int kFusedKey = 100;     // mode key of a fused observer group: kFusedKey + NRL_FUSE_* flags

int launch_cell_tiles(Ctx &c, int mode_key, const NrlCellTileParams &p, int grid, size_t smem, cudaStream_t st) {
  switch (mode_key) {
    case kFusedKey + 0: return launch_cell_fused<0>(c, p, grid, smem, st);
    case kFusedKey + 1: return launch_cell_fused<1>(c, p, grid, smem, st);
    case kFusedKey + 2: return launch_cell_fused<2>(c, p, grid, smem, st);
    case kFusedKey + 3: return launch_cell_fused<3>(c, p, grid, smem, st);
    case kFusedKey + 6: return launch_cell_fused<6>(c, p, grid, smem, st);
    case kFusedKey + 7: return launch_cell_fused<7>(c, p, grid, smem, st);
    case kFusedKey + 10: return launch_cell_fused<10>(c, p, grid, smem, st);
    case kFusedKey + 11: return launch_cell_fused<11>(c, p, grid, smem, st);
    case kFusedKey + 14: return launch_cell_fused<14>(c, p, grid, smem, st);
    case kFusedKey + 15: return launch_cell_fused<15>(c, p, grid, smem, st);
    case -1: return launch_cell_tiles_mode<NRL_MODE_VOXELIZE, NRL_MAX_CH>(c, p, grid, smem, st, 0);
    case NRL_MODE_VOXELIZE: return launch_cell_tiles_mode<NRL_MODE_VOXELIZE, 1>(c, p, grid, smem, st, 0);
    case NRL_OBS_EXISTENCE: return launch_cell_tiles_mode<NRL_OBS_EXISTENCE, 1>(c, p, grid, smem, st, 1);
    case NRL_OBS_DIRECT_COUNT: return launch_cell_tiles_mode<NRL_OBS_DIRECT_COUNT, 1>(c, p, grid, smem, st, 2);
    case NRL_OBS_DISTINCT_COUNT: return launch_cell_tiles_mode<NRL_OBS_DISTINCT_COUNT, 1>(c, p, grid, smem, st, 3);
    case NRL_OBS_MEAN_DISTANCE: return launch_cell_tiles_mode<NRL_OBS_MEAN_DISTANCE, 1>(c, p, grid, smem, st, 4);
    case NRL_OBS_CUMULATIVE_WEIGHT: return launch_cell_tiles_mode<NRL_OBS_CUMULATIVE_WEIGHT, 1>(c, p, grid, smem, st, 5);
    case NRL_OBS_DENSITY: return launch_cell_tiles_mode<NRL_OBS_DENSITY, 1>(c, p, grid, smem, st, 6);
    case NRL_OBS_ECCENTRICITY: return launch_cell_tiles_mode<NRL_OBS_ECCENTRICITY, 1>(c, p, grid, smem, st, 8);
    default: return fail(NRL_ERR_OBSERVABLE, "The observable type has no cell-list kernel");
  }
}

// Cell grid of one crop: what its cell-list tasks need.  Cells are sized so that the search box of
// a tile (7 voxel steps + twice the reach) spans at most 5 cells per axis: <= 25 runs + the NaN bucket.
struct CropCells {
  bool any = false;
  float reach = 0.f, span = 0.f, ext[3] = {0.f, 0.f, 0.f};
  void add(const int *dims, float spacing, float cutoff) {
    const float r = cutoff * 1.001f + 1e-3f;
    any = true;
    reach = std::max(reach, r);
    span = std::max(span, 7.0f * spacing + 2.0f * r);
    ext[0] = std::max(ext[0], (float)(dims[2] - 1) * spacing);
    ext[1] = std::max(ext[1], (float)(dims[1] - 1) * spacing);
    ext[2] = std::max(ext[2], (float)(dims[0] - 1) * spacing);
  }
  // false: the grid would need more cells than the binning kernel can count in shared memory
  bool grid(NrlCellGrid &g, size_t max_cells) const {
    const float h = std::max(span / 3.9f, 1e-6f);
    const float pad = reach + 2e-3f;
    for (int zdiv = 3; zdiv >= 1; --zdiv) {
      const float hz = h / (float)zdiv;
      const double ncx = std::floor((ext[0] + 2.0 * pad) / h) + 1, ncy = std::floor((ext[1] + 2.0 * pad) / h) + 1,
                   ncz = std::floor((ext[2] + 2.0 * pad) / hz) + 1;
      if (ncx * ncy * ncz + 2 > (double)max_cells) continue;
      g.ox = g.oy = g.oz = -pad;
      g.inv_h = 1.0f / h;
      g.inv_hz = 1.0f / hz;
      g.ncx = (int)ncx; g.ncy = (int)ncy; g.ncz = (int)ncz;
      g.cs_off = 0;
      g.bx = ext[0]; g.by = ext[1]; g.bz = ext[2];
      g.reach2 = pad * pad;
      return true;
    }
    return false;
  }
};

constexpr size_t kMaxCells = 48 * 1024;     // counters of the binning kernel: 192 KB of shared memory

// bins `frame_list` (host) of `packed` into c.batoms / c.bcnt / c.cstart with the grids in `hgrids` (per frame
// descriptor, cs_off filled here); returns the device-side setup
int bin_frames(Ctx &c, const NrlFrames &packed, std::vector<NrlCellGrid> &hgrids, const std::vector<int> &frame_list,
               size_t atoms_capacity, int capacity, cudaStream_t st, CellSetup &out) {
  long long cs_total = 0;
  int max_ncell = 0;
  for (int fd : frame_list) {
    NrlCellGrid &g = hgrids[fd];
    const int ncell = g.ncx * g.ncy * g.ncz;
    g.cs_off = cs_total;
    cs_total += ncell + 2;
    max_ncell = std::max(max_ncell, ncell);
  }
  NRL_TRY(c.batoms.reserve(std::max<size_t>(atoms_capacity, 1) * sizeof(float4)));
  NRL_TRY(c.bcnt.reserve(hgrids.size() * sizeof(int)));
  NRL_TRY(c.cstart.reserve((size_t)cs_total * sizeof(int)));
  // grids | frame list in one blob, one upload
  const size_t grids_bytes = (hgrids.size() * sizeof(NrlCellGrid) + 15) & ~(size_t)15;
  std::vector<unsigned char> blob(grids_bytes + frame_list.size() * sizeof(int));
  memcpy(blob.data(), hgrids.data(), hgrids.size() * sizeof(NrlCellGrid));
  memcpy(blob.data() + grids_bytes, frame_list.data(), frame_list.size() * sizeof(int));
  NRL_TRY(c.grids.reserve(blob.size()));
  NRL_CUDA(cudaMemcpyAsync(c.grids.p, blob.data(), blob.size(), cudaMemcpyHostToDevice, st));
  const size_t smem = (size_t)(max_ncell + 2) * sizeof(int);
  if (!c.attr_set[kAttrBin]) {
    NRL_CUDA(cudaFuncSetAttribute(nrl_bin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(c.smem_optin - 2048)));
    c.attr_set[kAttrBin] = true;
  }
  NrlBinParams bp;
  bp.in = packed;
  bp.grids = c.grids.as<NrlCellGrid>();
  bp.frame_list = reinterpret_cast<const int *>((const char *)c.grids.p + grids_bytes);
  bp.atoms_out = c.batoms.as<float4>();
  bp.cnt_out = c.bcnt.as<int>();
  bp.cell_start = c.cstart.as<int>();
  bp.capacity = capacity;
  for (size_t f0 = 0; f0 < frame_list.size(); f0 += 65535u * 16u) {
    const unsigned n = (unsigned)std::min<size_t>(65535u * 16u, frame_list.size() - f0);
    NrlBinParams q = bp;
    q.frame_list = bp.frame_list + f0;
    {
      ScopedProf sp(c, st, CAT_CELL_BIN);
      nrl_bin_kernel<<<n, NRL_BIN_NT, smem, st>>>(q);
    }
    g_launches.fetch_add(1);
    NRL_CUDA(cudaGetLastError());
  }
  out.on = true;
  out.frames = NrlFrames{c.batoms.as<float4>(), packed.off, c.bcnt.as<int>(), packed.pbase};
  out.grids = c.grids.as<NrlCellGrid>();
  out.cell_start = c.cstart.as<int>();
  return NRL_OK;
}

int mode_key_of(const NrlTask &t) {
  if (t.mode == NRL_OBS_FUSED) return kFusedKey + t.fuse_flags;
  return (t.mode == NRL_MODE_VOXELIZE && t.n_w > 1) ? kVoxelizeMulti : t.mode;
}
// flags of a set of observables (R and COM imply WS: their results divide by the sum of weights)
int fuse_flags_of(const int *modes, int n) {
  int f = 0;
  for (int i = 0; i < n; ++i) {
    if (modes[i] == NRL_OBS_DISTINCT_COUNT) f |= NRL_FUSE_SEEN;
    if (modes[i] == NRL_OBS_CUMULATIVE_WEIGHT || modes[i] == NRL_OBS_DENSITY) f |= NRL_FUSE_WS;
    if (modes[i] == NRL_OBS_MEAN_DISTANCE) f |= NRL_FUSE_WS | NRL_FUSE_R;
    if (modes[i] == NRL_OBS_ECCENTRICITY) f |= NRL_FUSE_WS | NRL_FUSE_COM;
  }
  return f;
}

void fill_task_geometry(NrlTask &t, const int *dims, float spacing, float cutoff, float sigma) {
  t.nx = dims[2]; t.ny = dims[1]; t.nz = dims[0];
  t.ntx = (t.nx + NRL_TX - 1) / NRL_TX;
  t.nty = (t.ny + NRL_TY - 1) / NRL_TY;
  t.ntz = (t.nz + NRL_TZ - 1) / NRL_TZ;
  t.tx0 = 0;
  t.spacing = spacing;
  t.cutoff = cutoff;
  t.sigma = sigma;
  t.k2 = sigma != 0.0f ? (float)(-0.5 * 1.4426950408889634 / ((double)sigma * (double)sigma)) : 0.0f;
  t.sphere_volume = (4.0 / 3.0) * M_PI * (double)cutoff * (double)cutoff * (double)cutoff;
}

// restrict a task to the slab [x_begin, x_begin + x_count) of its slowest axis (validated by the caller; x_count 0: no-op)
void apply_slab(NrlTask &t, int x_begin, int x_count) {
  if (x_count <= 0) return;
  t.tx0 = x_begin / NRL_TX;
  t.ntx = (x_count + NRL_TX - 1) / NRL_TX;
}

int ensure_aux(Ctx &c) {
  if (c.ev_fork) return NRL_OK;
  for (int i = 0; i < Ctx::kAux; ++i) {
    NRL_CUDA(cudaStreamCreateWithFlags(&c.aux[i], cudaStreamNonBlocking));
    NRL_CUDA(cudaEventCreateWithFlags(&c.ev_join[i], cudaEventDisableTiming));
  }
  NRL_CUDA(cudaEventCreateWithFlags(&c.ev_fork, cudaEventDisableTiming));
  return NRL_OK;
}

// Runs `tasks` (device pointers already filled in) grouped by (kernel family, mode): one launch per
// group.  `use_cell[i]` sends task i to the cell-list kernel (over cells.frames), otherwise it runs
// on the scan kernel (over `frames`).  The launches are independent: the first group stays on `st`,
// the others are forked onto side streams and joined back, so that their CTAs fill the SMs
// together (a 32^3 task is only 64 tiles).  `before_voxelize` enqueues, on `st`, whatever the
// voxeliser groups depend on (the per-atom normalisation); those groups then stay on `st`, the
// observers do not wait for it.  max_atoms_per_frame bounds the per-CTA overflow candidate list.
template <typename PreVox>
int run_tasks(Ctx &c, std::vector<NrlTask> &tasks, const std::vector<char> &use_cell, const NrlFrames &frames,
              const CellSetup &cells, float *d_out, int max_atoms_per_frame, cudaStream_t st, PreVox before_voxelize) {
  if (tasks.empty()) return NRL_OK;
  // cell-list tasks whose reducer streams (no series) are launched apart: their CTAs need much less shared memory
  auto key_of = [&](int i) { return mode_key_of(tasks[i]) + (use_cell[i] ? (nrl_streaming_task(tasks[i]) ? 2000 : 1000) : 0); };
  std::vector<int> order(tasks.size());
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key_of(a) < key_of(b); });
  std::vector<NrlTask> sorted;
  std::vector<int> keys;
  sorted.reserve(tasks.size());
  for (int i : order) { sorted.push_back(tasks[i]); keys.push_back(key_of(i)); }

  // per-group ranges with their own tile prefix
  struct Range { int mode, begin, end, n_work, max_frames, grid, cand_cap, ser_cap, wcap; bool cell, stream;
                 size_t smem, cand_rec, cand_per_cta, ser_per_cta, cand_off, ser_off; };
  std::vector<Range> ranges;
  for (size_t i = 0; i < sorted.size();) {
    size_t j = i;
    int work = 0, maxf = 1;
    while (j < sorted.size() && keys[j] == keys[i]) {
      sorted[j].tile_begin = work;
      work += sorted[j].ntx * sorted[j].nty * sorted[j].ntz;
      maxf = std::max(maxf, sorted[j].n_frames * nrl_series_rows(sorted[j]));     // series rows a tile keeps
      ++j;
    }
    Range r{};
    r.mode = mode_key_of(sorted[i]); r.cell = keys[i] >= 500; r.stream = keys[i] >= 1500; r.begin = (int)i; r.end = (int)j; r.n_work = work; r.max_frames = maxf;
    ranges.push_back(r);
    i = j;
  }
  NRL_TRY(c.tasks.reserve(sorted.size() * sizeof(NrlTask)));
  NRL_CUDA(cudaMemcpyAsync(c.tasks.p, sorted.data(), sorted.size() * sizeof(NrlTask), cudaMemcpyHostToDevice, st));

  // launch geometry and overflow scratch of every range (disjoint slices: the ranges run concurrently)
  size_t cand_total = 0, ser_total = 0;
  for (Range &r : ranges) {
    const bool is_vox = r.mode == NRL_MODE_VOXELIZE || r.mode == kVoxelizeMulti;
    // shared memory: the candidate records plus the time series when it fits (else per-CTA global scratch)
    const size_t ser_need = r.stream ? 0 : (size_t)r.max_frames * NRL_TILE_VOX;
    r.ser_cap = ser_need <= (size_t)kSerCapFloats ? (int)ser_need : 0;
    if (r.cell) {
      // staged cells + per-warp records (16 B; voxeliser 32 B / 48 B) + series
      const bool fused_seen = r.mode >= kFusedKey && ((r.mode - kFusedKey) & NRL_FUSE_SEEN);
      const int rec4 = r.mode == kVoxelizeMulti ? 3 : (is_vox || fused_seen) ? 2 : 1;
      r.wcap = r.mode == kVoxelizeMulti ? 128 : (is_vox || fused_seen) ? 160 : 192;
      r.smem = (size_t)NRL_STAGE_CAP * sizeof(float4) + (size_t)NRL_NW * r.wcap * rec4 * sizeof(float4) + (size_t)r.ser_cap * sizeof(float);
      r.cand_per_cta = 0;
    } else {
      // candidate records held in shared memory: 512 atoms, 384 of the voxeliser's 48-byte records
      r.cand_rec = (r.mode == kVoxelizeMulti ? 4 : is_vox ? 3 : 1) * sizeof(float4);
      r.cand_cap = is_vox ? 384 : kCandCap;
      r.smem = r.cand_cap * r.cand_rec + (size_t)r.ser_cap * sizeof(float) + (size_t)NRL_NW * r.cand_cap * sizeof(unsigned short);
      // overflow scratch (candidate lists longer than kCandCap, series longer than kSerCapFloats)
      r.cand_per_cta = max_atoms_per_frame + 4 > r.cand_cap ? (size_t)max_atoms_per_frame : 0;
    }
    const int ctas_per_sm = std::max(1, (int)(c.smem_optin / (r.smem + 1024)));
    // a tile is tens of microseconds of work: up to a few waves one CTA per tile lets the block scheduler
    // balance the SMs; beyond that the CTAs stride over the tiles
    const int slots = c.sms * std::min(ctas_per_sm, 12);
    r.grid = r.n_work <= 8 * slots ? r.n_work : slots;
    r.ser_per_cta = r.ser_cap == 0 ? ser_need : 0;
    const size_t per_cta = r.cand_per_cta * r.cand_rec + r.ser_per_cta * sizeof(float);
    if (per_cta > 0) {
      const int max_grid = (int)std::max<size_t>((size_t)c.sms, kScratchBudget / ranges.size() / per_cta);
      r.grid = std::max(1, std::min(r.grid, max_grid));
    }
    r.cand_off = cand_total;
    r.ser_off = ser_total;
    cand_total += (r.cand_per_cta * r.cand_rec * r.grid + 255) & ~(size_t)255;
    ser_total += (r.ser_per_cta * sizeof(float) * r.grid + 255) & ~(size_t)255;
  }
  NRL_TRY(c.cand_scr.reserve(cand_total));
  NRL_TRY(c.ser_scr.reserve(ser_total));

  const bool fork = ranges.size() > 1 && !g_no_fork;
  if (fork) {
    NRL_TRY(ensure_aux(c));
    NRL_CUDA(cudaEventRecord(c.ev_fork, st));
  }
  bool aux_used[Ctx::kAux] = {false, false, false};
  int next_aux = 0;
  bool vox_ready = false;
  for (const Range &r : ranges) {
    const bool is_vox = r.mode == NRL_MODE_VOXELIZE || r.mode == kVoxelizeMulti;
    cudaStream_t rs = st;
    if (is_vox) {
      if (!vox_ready) { NRL_TRY(before_voxelize()); vox_ready = true; }
    } else if (fork) {
      const int k = next_aux++ % Ctx::kAux;
      rs = c.aux[k];
      if (!aux_used[k]) { NRL_CUDA(cudaStreamWaitEvent(rs, c.ev_fork, 0)); aux_used[k] = true; }
    }
    if (r.cell) {
      NrlCellTileParams p;
      p.tasks = c.tasks.as<NrlTask>() + r.begin;
      p.n_tasks = r.end - r.begin;
      p.n_work = r.n_work;
      p.frames = cells.frames;
      p.grids = cells.grids;
      p.cell_start = cells.cell_start;
      p.out = d_out;
      p.ser_scratch = reinterpret_cast<float *>(c.ser_scr.as<char>() + r.ser_off);
      p.ser_scratch_stride = (long long)r.ser_per_cta;
      p.ser_cap = r.ser_cap;
      p.wcap = r.wcap;
      NRL_TRY(launch_cell_tiles(c, r.mode, p, r.grid, r.smem, rs));
    } else {
      NrlTileParams p;
      p.tasks = c.tasks.as<NrlTask>() + r.begin;
      p.n_tasks = r.end - r.begin;
      p.n_work = r.n_work;
      p.frames = frames;
      p.out = d_out;
      p.cand_scratch = reinterpret_cast<float4 *>(c.cand_scr.as<char>() + r.cand_off);
      p.cand_scratch_stride = (long long)(r.cand_per_cta * r.cand_rec / sizeof(float4));
      p.ser_scratch = reinterpret_cast<float *>(c.ser_scr.as<char>() + r.ser_off);
      p.ser_scratch_stride = (long long)r.ser_per_cta;
      p.cand_cap = r.cand_cap;
      p.ser_cap = r.ser_cap;
      NRL_TRY(launch_tiles(c, r.mode, p, r.grid, r.smem, rs));
    }
  }
  for (int k = 0; k < Ctx::kAux; ++k)
    if (aux_used[k]) {
      NRL_CUDA(cudaEventRecord(c.ev_join[k], c.aux[k]));
      NRL_CUDA(cudaStreamWaitEvent(st, c.ev_join[k], 0));
    }
  return NRL_OK;
}

int launch_norm(Ctx &c, const NrlFrames &frames, const int *d_frame_list, int n_list, int capacity, const NrlTask &geom,
                float *d_inv, cudaStream_t st) {
  if (n_list <= 0 || capacity <= 0) return NRL_OK;
  NrlNormParams np;
  np.frames = frames;
  np.frame_list = d_frame_list;
  np.nx = geom.nx; np.ny = geom.ny; np.nz = geom.nz;
  np.spacing = geom.spacing; np.cutoff = geom.cutoff; np.k2 = geom.k2;
  np.inv_norm = d_inv;
  const int warps_per_cta = 4;
  for (int f0 = 0; f0 < n_list; f0 += 65535) {
    const int ny = std::min(65535, n_list - f0);
    NrlNormParams q = np;
    q.frame_list = d_frame_list + f0;
    // warps stride over the atoms of their frame: at least 48 CTAs x 4 warps per frame, and for a launch of few
    // frames enough CTAs to put ~32 warps on every SM (a single-frame call is latency bound otherwise)
    const int want = std::max(48, (c.sms * 32 / warps_per_cta + ny - 1) / ny);
    dim3 grid((unsigned)std::max(1, std::min((capacity + warps_per_cta - 1) / warps_per_cta, want)), (unsigned)ny);
    {
      ScopedProf sp(c, st, CAT_ATOM_NORM);
      nrl_atom_norm_kernel<<<grid, warps_per_cta * 32, 0, st>>>(q, capacity);
    }
    g_launches.fetch_add(1);
    NRL_CUDA(cudaGetLastError());
  }
  return NRL_OK;
}

// value order of every frame's weights for distinct_count: equal values adjacent, NaNs last
void host_value_order(const float *w, int F, int N, std::vector<int> &perm) {
  perm.resize((size_t)F * N);
  for (int f = 0; f < F; ++f) {
    int *p = perm.data() + (size_t)f * N;
    const float *wf = w + (size_t)f * N;
    std::iota(p, p + N, 0);
    std::stable_sort(p, p + N, [wf](int a, int b) {
      const float x = wf[a], y = wf[b];
      const bool nx = x != x, ny = y != y;
      if (nx || ny) return !nx && ny;
      return x < y;
    });
  }
}

// ---- shared implementation of the four "reference layout" operations --------------------
struct LegacyArgs {
  const float *coords, *weights;
  int F, N;
  const int *dims;
  float spacing, cutoff, sigma;
  int mode, agg;       // agg 0: single frame
  float *out;
  bool on_device;      // pointers are device pointers, no sync
  cudaStream_t user_stream;
};

int legacy_run(const LegacyArgs &a) {
  long long G = 0;
  NRL_TRY(check_dims(a.dims, &G));
  if (a.F <= 0 || a.N < 0) return fail(NRL_ERR_SHAPE, "frame / atom counts must be positive");
  if (!a.out) return fail(NRL_ERR_ARGUMENT, "out is NULL");
  if (a.N > 0 && (!a.coords || !a.weights)) return fail(NRL_ERR_ARGUMENT, "coords / weights is NULL");
  Ctx *cp = nullptr;
  NRL_TRY(get_ctx(&cp));
  Ctx &c = *cp;
  std::lock_guard<std::mutex> lk(c.mu);
  cudaStream_t st = a.on_device ? a.user_stream : c.work_stream();
  WorkspaceTurn turn(c, st);
  const int nF = std::min(a.F, NRL_MAX_FRAMES);   // gpuutils.cu:88, voxelize.cu:403, marching_observers.cu:490
  const size_t n_atoms = (size_t)nF * (size_t)a.N;

  float *d_out = a.out;
  if (!a.on_device) { NRL_TRY(c.out.reserve((size_t)G * sizeof(float))); d_out = c.out.as<float>(); }

  if (a.mode == NRL_MODE_VOXELIZE && a.sigma == 0.0f) {
    // gaussian_map_device returns 0 for sigma == 0 (gpuutils.cuh:312-313): every sum is 0, nothing is deposited
    NRL_CUDA(cudaMemsetAsync(d_out, 0, (size_t)G * sizeof(float), st));
  } else {
    const float *d_coords = a.coords, *d_weights = a.weights;
    std::vector<int> perm;
    if (!a.on_device && n_atoms > 0) {
      NRL_TRY(c.coords.reserve(n_atoms * 3 * sizeof(float)));
      NRL_TRY(c.weights.reserve(n_atoms * sizeof(float)));
      NRL_CUDA(cudaMemcpyAsync(c.coords.p, a.coords, n_atoms * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
      NRL_CUDA(cudaMemcpyAsync(c.weights.p, a.weights, n_atoms * sizeof(float), cudaMemcpyHostToDevice, st));
      d_coords = c.coords.as<float>();
      d_weights = c.weights.as<float>();
    }
    NrlCellGrid cell_grid;
    bool grid_ok = false;
    {
      CropCells cc;
      cc.add(a.dims, a.spacing, a.cutoff);
      grid_ok = cc.grid(cell_grid, kMaxCells);
    }
    const int *d_perm = nullptr;
    const float *d_ranks = nullptr;       // distinct_count over <= 32 distinct values: value ranks for the cell-list kernel
    if (a.mode == NRL_OBS_DISTINCT_COUNT && n_atoms > 0) {
      std::vector<float> hw;
      const float *w_host = a.weights;
      if (a.on_device) {
        hw.resize(n_atoms);
        NRL_CUDA(cudaMemcpyAsync(hw.data(), a.weights, n_atoms * sizeof(float), cudaMemcpyDeviceToHost, st));
        NRL_CUDA(cudaStreamSynchronize(st));
        w_host = hw.data();
      }
      // few distinct values (atomic numbers, flags ...): one pass collects them, their sorted position is the rank
      if (!g_force_scan && grid_ok) {
        // 64-slot open-addressing set keyed on the value's bits (-0.0 folded onto +0.0, as float == does)
        uint32_t keys[64];
        int slot_val[64], nv = 0;
        float vals[32];
        bool used[64] = {false}, few = true;
        auto find = [&](float v, bool insert) -> int {
          if (v == 0.0f) v = 0.0f;
          uint32_t bits;
          memcpy(&bits, &v, 4);
          uint32_t h = (bits * 2654435761u) >> 26;
          while (used[h] && keys[h] != bits) h = (h + 1) & 63;
          if (!used[h]) {
            if (!insert || nv == 32) return -1;
            used[h] = true; keys[h] = bits; slot_val[h] = nv; vals[nv++] = v;
          }
          return slot_val[h];
        };
        for (size_t i = 0; i < n_atoms && few; ++i) {
          const float v = w_host[i];
          if (v != v || find(v, true) < 0) few = false;       // every NaN is a value of its own: leave it to the scan kernel
        }
        if (few) {
          int order[32], rank_of[32];
          std::iota(order, order + nv, 0);
          std::sort(order, order + nv, [&](int x, int y) { return vals[x] < vals[y]; });
          for (int r = 0; r < nv; ++r) rank_of[order[r]] = r;
          std::vector<float> ranks(n_atoms);
          for (size_t i = 0; i < n_atoms; ++i) {
            const float v = w_host[i];
            ranks[i] = (v == NRL_PLACEHOLDER) ? -1.0f : (float)rank_of[find(v, false)];            // marching_observers.cu:118-138
          }
          NRL_TRY(c.perm.reserve(n_atoms * sizeof(float)));
          NRL_CUDA(cudaMemcpyAsync(c.perm.p, ranks.data(), n_atoms * sizeof(float), cudaMemcpyHostToDevice, st));
          d_ranks = c.perm.as<float>();
        }
      }
    }
    if (a.mode == NRL_OBS_DISTINCT_COUNT && n_atoms > 0 && !d_ranks) {
      std::vector<float> hw;
      const float *w_host = a.weights;
      if (a.on_device) {
        hw.resize(n_atoms);
        NRL_CUDA(cudaMemcpyAsync(hw.data(), a.weights, n_atoms * sizeof(float), cudaMemcpyDeviceToHost, st));
        NRL_CUDA(cudaStreamSynchronize(st));
        w_host = hw.data();
      }
      host_value_order(w_host, nF, a.N, perm);
      NRL_TRY(c.perm.reserve(n_atoms * sizeof(int)));
      NRL_CUDA(cudaMemcpyAsync(c.perm.p, perm.data(), n_atoms * sizeof(int), cudaMemcpyHostToDevice, st));
      d_perm = c.perm.as<int>();
    }
    NRL_TRY(c.atoms.reserve(std::max<size_t>(n_atoms, 1) * sizeof(float4)));
    NRL_TRY(c.off.reserve((size_t)nF * sizeof(long long)));
    NRL_TRY(c.cnt.reserve((size_t)nF * sizeof(int)));
    NRL_TRY(c.pbase.reserve((size_t)nF * sizeof(int)));
    {
      const long long total = (long long)n_atoms;
      const int grid = (int)std::min<long long>(std::max<long long>((std::max<long long>(total, nF) + 255) / 256, 1), 4096);
      {
        ScopedProf sp(c, st, CAT_PACK_LEGACY);
        nrl_pack_legacy_kernel<<<grid, 256, 0, st>>>(d_coords, d_weights, d_perm, nF, a.N,
                                                      (a.mode == NRL_OBS_DISTINCT_COUNT && !d_ranks) ? 1 : 0, c.atoms.as<float4>(),
                                                      c.off.as<long long>(), c.cnt.as<int>(), c.pbase.as<int>());
      }
      g_launches.fetch_add(1);
      NRL_CUDA(cudaGetLastError());
    }
    NrlFrames frames{c.atoms.as<float4>(), c.off.as<long long>(), c.cnt.as<int>(), c.pbase.as<int>()};
    NrlTask t;
    memset(&t, 0, sizeof t);
    fill_task_geometry(t, a.dims, a.spacing, a.cutoff, a.sigma);
    t.n_frames = nF;
    t.n_ch = 1;
    t.n_w = 1;
    t.ch_agg[0] = a.agg;
    t.ch_w[0] = 0;
    t.ch_out[0] = 0;
    t.mode = a.mode;
    t.frame0 = 0;
    t.wtab[0] = d_ranks ? d_ranks : d_weights;
    t.inv_norm = nullptr;
    // cell lists for every mode that has a cell-list kernel (g_force_scan: tests / A-B timing)
    CellSetup cells;
    // (a single short frame is cheaper on the scan kernel: nothing to amortise the binning launch over;
    //  rank-based distinct_count exists on the cell-list kernel only)
    if ((cell_mode(a.mode) || d_ranks) && !g_force_scan && !index_order_matters(a.mode, a.agg) && n_atoms > 0 &&
        (d_ranks || !(nF == 1 && a.N <= 4096))) {
      if (grid_ok) {
        std::vector<NrlCellGrid> hgrids((size_t)nF, cell_grid);
        std::vector<int> fl(nF);
        std::iota(fl.begin(), fl.end(), 0);
        NRL_TRY(bin_frames(c, frames, hgrids, fl, n_atoms, a.N, st, cells));
      }
    }
    const NrlFrames &vf = cells.on ? cells.frames : frames;
    if (a.mode == NRL_MODE_VOXELIZE) {
      NRL_TRY(c.inv.reserve(std::max<size_t>(n_atoms, 1) * sizeof(float)));
      NRL_TRY(c.flist.reserve((size_t)nF * sizeof(int)));
      std::vector<int> fl(nF);
      std::iota(fl.begin(), fl.end(), 0);
      NRL_CUDA(cudaMemcpyAsync(c.flist.p, fl.data(), (size_t)nF * sizeof(int), cudaMemcpyHostToDevice, st));
      t.inv_norm = c.inv.as<float>();
    }
    std::vector<NrlTask> tasks{t};
    std::vector<char> use_cell{(char)(cells.on ? 1 : 0)};
    NRL_TRY(run_tasks(c, tasks, use_cell, frames, cells, d_out, a.N, st, [&]() -> int {
      return launch_norm(c, vf, c.flist.as<int>(), nF, a.N, t, c.inv.as<float>(), st);
    }));
  }
  if (!a.on_device) {
    NRL_CUDA(cudaMemcpyAsync(a.out, d_out, (size_t)G * sizeof(float), cudaMemcpyDeviceToHost, st));
    NRL_CUDA(cudaStreamSynchronize(st));
    prof_collect(c);
  }
  return NRL_OK;
}

// cuTensorMapEncodeTiled through the runtime (the library does not link libcuda); nullptr when the driver lacks it
typedef CUresult (*nrl_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
nrl_encode_tiled_fn tensor_map_encoder() {
  static nrl_encode_tiled_fn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      p = nullptr;
    }
    return reinterpret_cast<nrl_encode_tiled_fn>(p);
  }();
  return fn;
}

int aggregate_impl(const float *series, int F, long long G, int agg, float *out, bool on_device, cudaStream_t user) {
  if (F <= 0 || G <= 0) return fail(NRL_ERR_SHAPE, "series must be [n_frames > 0, n_voxels > 0]");
  if (!series || !out) return fail(NRL_ERR_ARGUMENT, "series / out is NULL");
  Ctx *cp = nullptr;
  NRL_TRY(get_ctx(&cp));
  Ctx &c = *cp;
  std::lock_guard<std::mutex> lk(c.mu);
  cudaStream_t st = on_device ? user : c.work_stream();
  WorkspaceTurn turn(c, st);
  const int n = std::min(F, NRL_MAX_FRAMES);                 // gpuutils.cu:88
  const float *d_in = series;
  float *d_out = out;
  if (!on_device) {
    NRL_TRY(c.series.reserve((size_t)n * G * sizeof(float)));
    NRL_TRY(c.out.reserve((size_t)G * sizeof(float)));
    NRL_CUDA(cudaMemcpyAsync(c.series.p, series, (size_t)n * G * sizeof(float), cudaMemcpyHostToDevice, st));
    d_in = c.series.as<float>();
    d_out = c.out.as<float>();
  }
  if (!valid_agg(agg)) {
    NRL_CUDA(cudaMemsetAsync(d_out, 0, (size_t)G * sizeof(float), st));   // gpuutils.cu:73-76
  } else if (agg == 1 || agg == 5 || agg == 6 || agg == 8) {
    // recurrences over the frames: streamed straight out of HBM, nothing staged
    const bool vec = (G % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_in) & 15) == 0) && ((reinterpret_cast<uintptr_t>(d_out) & 15) == 0);
    const long long threads = vec ? G / 4 : G;
    const long long grid = (threads + NRL_AGG_STREAM_NT - 1) / NRL_AGG_STREAM_NT;
    {
      ScopedProf sp(c, st, CAT_AGGREGATE);
      if (vec) nrl_aggregate_stream_kernel<4><<<(unsigned)grid, NRL_AGG_STREAM_NT, 0, st>>>(d_in, n, G, agg, d_out);
      else nrl_aggregate_stream_kernel<1><<<(unsigned)grid, NRL_AGG_STREAM_NT, 0, st>>>(d_in, n, G, agg, d_out);
    }
    g_launches.fetch_add(1);
    NRL_CUDA(cudaGetLastError());
  } else if ((agg == 2 || agg == 4 || agg == 7) && n <= 64) {
    // two passes over a short series: the series lives in registers (measured: beyond 64 frames the register
    // footprint leaves one CTA per SM and the shared-memory form below is faster)
    const long long grid = (G + 255) / 256;
    {
      ScopedProf sp(c, st, CAT_AGGREGATE);
      if (n <= 32) nrl_aggregate_reg_kernel<32><<<(unsigned)grid, 256, 0, st>>>(d_in, n, G, agg, d_out);
      else nrl_aggregate_reg_kernel<64><<<(unsigned)grid, 256, 0, st>>>(d_in, n, G, agg, d_out);
    }
    g_launches.fetch_add(1);
    NRL_CUDA(cudaGetLastError());
  } else if (n > 64 && agg != 3 && (G % 4 == 0) && G < (1ll << 31) && ((reinterpret_cast<uintptr_t>(d_in) & 15) == 0) &&
             g_agg_pipe_stages.load() >= 0 && tensor_map_encoder() != nullptr) {
    // long series, two-pass reducers (the median is ALU-bound and keeps the one-shot kernel, 4 CTAs per SM): persistent
    // CTAs, ring of staged tiles, one tiled TMA copy per tile (the copies of the next tiles run under the reduction)
    const size_t budget = c.smem_optin - 2048 - 128;
    const int n_boxes = (n + 255) / 256;
    const int rows = (n + n_boxes - 1) / n_boxes;
    const size_t tile_rows = (size_t)rows * n_boxes;
    int W = g_agg_pipe_w.load();
    if (W <= 0) {
      // measured on B200 (scripts/k3_pipe_probe.py): the widest tile that fits twice while it is >= 128 voxels wide,
      // else the widest that fits once (a reduction keeps W of the 256 threads busy: width beats ring depth)
      W = 256;
      while (W > 128 && tile_rows * W * sizeof(float) * 2 > budget) W >>= 1;
      while (W > 32 && tile_rows * W * sizeof(float) > budget) W >>= 1;
    }
    W = std::max(32, std::min(256, W & ~31));
    int stages = g_agg_pipe_stages.load();
    const int fit = (int)std::min<size_t>(NRL_AGG_PIPE_MAX_STAGES, budget / (tile_rows * W * sizeof(float)));
    if (fit < 1) return fail(NRL_ERR_FRAMES, "series too long for shared memory");
    stages = stages <= 0 ? fit : std::min(stages, fit);
    const size_t smem = (size_t)stages * tile_rows * W * sizeof(float) + 128;
    CUtensorMap map;
    {
      const cuuint64_t dims[2] = {(cuuint64_t)G, (cuuint64_t)n};
      const cuuint64_t strides[1] = {(cuuint64_t)G * sizeof(float)};
      const cuuint32_t box[2] = {(cuuint32_t)W, (cuuint32_t)rows};
      const cuuint32_t estr[2] = {1, 1};
      const CUresult r = tensor_map_encoder()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(d_in), dims, strides, box, estr,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(NRL_ERR_CUDA, "cuTensorMapEncodeTiled failed");
    }
    if (!c.attr_set[kAttrAggPipe]) {
      NRL_CUDA(cudaFuncSetAttribute(nrl_aggregate_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(budget + 128)));
      c.attr_set[kAttrAggPipe] = true;
    }
    const long long n_tiles = (G + W - 1) / W;
    const long long grid = std::min<long long>(n_tiles, c.sms);
    {
      ScopedProf sp(c, st, CAT_AGGREGATE);
      nrl_aggregate_pipe_kernel<<<(unsigned)grid, NRL_AGG_PIPE_NT, smem, st>>>(map, n, G, (agg == 7 && g_agg_entropy_legacy.load()) ? NRL_AGG_LEGACY_ENTROPY : agg, W, stages, n_tiles, rows, n_boxes, d_out);
    }
    g_launches.fetch_add(1);
    NRL_CUDA(cudaGetLastError());
  } else {
    // widest tile that still lets two CTAs share an SM
    int W = 512;
    while (W > 64 && (size_t)n * W * sizeof(float) > 56 * 1024) W >>= 1;   // four CTAs per SM: copies of some overlap the reduction of others   // four CTAs per SM: copies of some overlap the reduction of others
    const size_t smem = (size_t)n * W * sizeof(float);
    if (smem > c.smem_optin - 1024) return fail(NRL_ERR_FRAMES, "series too long for shared memory");
    if (!c.attr_set[kAttrAggStaged]) {
      NRL_CUDA(cudaFuncSetAttribute(nrl_aggregate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(c.smem_optin - 1024)));
      c.attr_set[kAttrAggStaged] = true;
    }
    const int bulk = (G % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_in) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(d_out) & 15) == 0);
    const long long grid = (G + W - 1) / W;
    {
      ScopedProf sp(c, st, CAT_AGGREGATE);
      nrl_aggregate_kernel<<<(unsigned)grid, NRL_AGG_NT, smem, st>>>(d_in, n, G, (agg == 7 && g_agg_entropy_legacy.load()) ? NRL_AGG_LEGACY_ENTROPY : agg, W, bulk, d_out);
    }
    g_launches.fetch_add(1);
    NRL_CUDA(cudaGetLastError());
  }
  if (!on_device) {
    NRL_CUDA(cudaMemcpyAsync(out, d_out, (size_t)G * sizeof(float), cudaMemcpyDeviceToHost, st));
    NRL_CUDA(cudaStreamSynchronize(st));
    prof_collect(c);
  }
  return NRL_OK;
}

}  // namespace

extern "C" {

const char *nrl_last_error(void) { return g_err.c_str(); }
const char *nrl_version(void) { return "nearl_b200 0.1 (sm_100a)"; }
long long nrl_kernel_launches(void) { return g_launches.load(); }

int nrl_device_count(int *out) {
  if (!out) return fail(NRL_ERR_ARGUMENT, "out is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
  *out = n;
  return NRL_OK;
}
int nrl_set_device(int device) { NRL_CUDA(cudaSetDevice(device)); return NRL_OK; }
int nrl_synchronize(void) {
  Ctx *c = nullptr;
  NRL_TRY(get_ctx(&c));
  std::lock_guard<std::mutex> lk(c->mu);
  NRL_CUDA(cudaStreamSynchronize(c->work_stream()));
  prof_collect(*c);
  return NRL_OK;
}
int nrl_lane_synchronize(int lane) {
  Ctx *c = nullptr;
  NRL_TRY(get_ctx(&c, lane));
  std::lock_guard<std::mutex> lk(c->mu);
  NRL_CUDA(cudaStreamSynchronize(c->work_stream()));
  prof_collect(*c);
  return NRL_OK;
}
int nrl_set_stream(void *stream, int enable) {
  Ctx *c = nullptr;
  NRL_TRY(get_ctx(&c));
  std::lock_guard<std::mutex> lk(c->mu);
  NRL_CUDA(cudaStreamSynchronize(c->work_stream()));
  c->user_stream = (cudaStream_t)stream;
  c->use_user_stream = enable != 0;
  return NRL_OK;
}
int nrl_superpose(void *xyz, int xyz_is_f64, int n_frames, int n_atoms, const int *fit_index, int n_fit, int ref_frame,
                  int on_device, float *rmsd, void *stream) {
  if (!xyz || !fit_index) return fail(NRL_ERR_ARGUMENT, "xyz / fit_index is NULL");
  if (n_frames <= 0 || n_atoms <= 0 || n_fit <= 0) return fail(NRL_ERR_SHAPE, "empty trajectory or fit set");
  if (ref_frame < 0 || ref_frame >= n_frames) return fail(NRL_ERR_ARGUMENT, "reference frame outside the trajectory");
  for (int i = 0; i < n_fit; ++i)
    if (fit_index[i] < 0 || fit_index[i] >= n_atoms) return fail(NRL_ERR_ARGUMENT, "fit index outside the topology");
  Ctx *cp = nullptr;
  NRL_TRY(get_ctx(&cp));
  Ctx &c = *cp;
  std::lock_guard<std::mutex> lk(c.mu);
  cudaStream_t st = on_device ? (cudaStream_t)stream : c.work_stream();
  WorkspaceTurn turn(c, st);
  const size_t esz = xyz_is_f64 ? sizeof(double) : sizeof(float);
  const size_t bytes = (size_t)n_frames * n_atoms * 3 * esz;
  void *d_xyz = xyz;
  if (!on_device) {
    NRL_TRY(c.xyz.reserve(bytes));
    NRL_CUDA(cudaMemcpyAsync(c.xyz.p, xyz, bytes, cudaMemcpyHostToDevice, st));
    d_xyz = c.xyz.p;
  }
  NRL_TRY(c.sup_idx.reserve((size_t)n_fit * sizeof(int)));
  NRL_TRY(c.sup_ref.reserve((size_t)(4 + 3 * (size_t)n_fit) * sizeof(double)));
  NRL_TRY(c.sup_rmsd.reserve((size_t)n_frames * sizeof(float)));
  NRL_CUDA(cudaMemcpyAsync(c.sup_idx.p, fit_index, (size_t)n_fit * sizeof(int), cudaMemcpyHostToDevice, st));
  {
    ScopedProf sp(c, st, CAT_SUPERPOSE);
    NRL_TRY(c.sup_xform.reserve((size_t)n_frames * 12 * sizeof(double)));
    // apply: enough CTAs per frame for ~8 resident CTAs per SM over all frames, at most one per 1024 atoms
    const int per_frame = std::max(1, std::min((n_atoms / 4 + NRL_SUPA_NT - 1) / NRL_SUPA_NT, std::max(1, (c.sms * 8 + n_frames - 1) / n_frames)));
    const dim3 agrid((unsigned)per_frame, (unsigned)n_frames);
    if (xyz_is_f64) {
      nrl_superpose_ref_kernel<double><<<1, NRL_SUP_NT, 0, st>>>((const double *)d_xyz, n_atoms, c.sup_idx.as<int>(), n_fit, ref_frame, c.sup_ref.as<double>());
      nrl_superpose_kernel<double><<<n_frames, NRL_SUP_NT, 0, st>>>((double *)d_xyz, n_atoms, c.sup_idx.as<int>(), n_fit, c.sup_ref.as<double>(), ref_frame, c.sup_rmsd.as<float>(), c.sup_xform.as<double>());
      nrl_superpose_apply_kernel<double><<<agrid, NRL_SUPA_NT, 0, st>>>((double *)d_xyz, n_atoms, c.sup_xform.as<double>(), c.sup_ref.as<double>(), ref_frame);
    } else {
      nrl_superpose_ref_kernel<float><<<1, NRL_SUP_NT, 0, st>>>((const float *)d_xyz, n_atoms, c.sup_idx.as<int>(), n_fit, ref_frame, c.sup_ref.as<double>());
      nrl_superpose_kernel<float><<<n_frames, NRL_SUP_NT, 0, st>>>((float *)d_xyz, n_atoms, c.sup_idx.as<int>(), n_fit, c.sup_ref.as<double>(), ref_frame, c.sup_rmsd.as<float>(), c.sup_xform.as<double>());
      nrl_superpose_apply_kernel<float><<<agrid, NRL_SUPA_NT, 0, st>>>((float *)d_xyz, n_atoms, c.sup_xform.as<double>(), c.sup_ref.as<double>(), ref_frame);
    }
  }
  g_launches.fetch_add(3);
  NRL_CUDA(cudaGetLastError());
  if (rmsd) {
    if (on_device) NRL_CUDA(cudaMemcpyAsync(rmsd, c.sup_rmsd.p, (size_t)n_frames * sizeof(float), cudaMemcpyDeviceToDevice, st));
    else NRL_CUDA(cudaMemcpyAsync(rmsd, c.sup_rmsd.p, (size_t)n_frames * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  if (!on_device) {
    NRL_CUDA(cudaMemcpyAsync(xyz, d_xyz, bytes, cudaMemcpyDeviceToHost, st));
    NRL_CUDA(cudaStreamSynchronize(st));
    prof_collect(c);
  }
  return NRL_OK;
}

int nrl_cell_list_fits(const int *dims, float spacing, float cutoff, int *out) {
  long long G = 0;
  NRL_TRY(check_dims(dims, &G));
  if (!out) return fail(NRL_ERR_ARGUMENT, "out is NULL");
  CropCells cc;
  cc.add(dims, spacing, cutoff);
  NrlCellGrid g;
  *out = (!g_force_scan && cc.grid(g, kMaxCells)) ? 1 : 0;
  return NRL_OK;
}
int nrl_set_option(const char *name, int value) {
  if (!name) return fail(NRL_ERR_ARGUMENT, "option name is NULL");
  if (strcmp(name, "force_scan") == 0) { g_force_scan.store(value != 0); return NRL_OK; }
  if (strcmp(name, "no_fuse") == 0) { g_no_fuse.store(value != 0); return NRL_OK; }
  if (strcmp(name, "crop_legacy") == 0) { g_crop_legacy.store(value != 0); return NRL_OK; }
  if (strcmp(name, "no_fork") == 0) { g_no_fork.store(value != 0); return NRL_OK; }
  if (strcmp(name, "crop_stream_min") == 0) { g_crop_stream_min.store(value > 0 ? value : 2048); return NRL_OK; }
  if (strcmp(name, "agg_entropy_legacy") == 0) { g_agg_entropy_legacy.store(value != 0); return NRL_OK; }
  if (strcmp(name, "agg_pipe_w") == 0) { g_agg_pipe_w.store((int)value); return NRL_OK; }
  if (strcmp(name, "agg_pipe_stages") == 0) { g_agg_pipe_stages.store((int)value); return NRL_OK; }
  return fail(NRL_ERR_ARGUMENT, std::string("unknown option ") + name);
}
// every lane of the current device (lanes that were never used are created: a context is a stream and empty arenas)
int nrl_profile(int enable) {
  for (int lane = 0; lane < NRL_MAX_LANES; ++lane) {
    Ctx *c = nullptr;
    NRL_TRY(get_ctx(&c, lane));
    std::lock_guard<std::mutex> lk(c->mu);
    c->prof_on = enable != 0;
  }
  return NRL_OK;
}
int nrl_profile_read(double *ms, long long *counts, int reset) {
  for (int i = 0; i < NRL_PROFILE_CATEGORIES; ++i) {
    if (ms) ms[i] = 0;
    if (counts) counts[i] = 0;
  }
  for (int lane = 0; lane < NRL_MAX_LANES; ++lane) {
    Ctx *c = nullptr;
    NRL_TRY(get_ctx(&c, lane));
    std::lock_guard<std::mutex> lk(c->mu);
    NRL_CUDA(cudaStreamSynchronize(c->work_stream()));
    prof_collect(*c);
    for (int i = 0; i < NRL_PROFILE_CATEGORIES; ++i) {
      if (ms) ms[i] += c->prof_ms[i];
      if (counts) counts[i] += c->prof_cnt[i];
      if (reset) { c->prof_ms[i] = 0; c->prof_cnt[i] = 0; }
    }
  }
  return NRL_OK;
}
const char *nrl_profile_name(int category) {
  return (category >= 0 && category < NRL_PROFILE_CATEGORIES) ? kCatNames[category] : "";
}
int nrl_host_register(void *ptr, long long bytes) {
  if (!ptr || bytes <= 0) return fail(NRL_ERR_ARGUMENT, "nothing to register");
  NRL_CUDA(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault));
  return NRL_OK;
}
int nrl_host_unregister(void *ptr) { NRL_CUDA(cudaHostUnregister(ptr)); return NRL_OK; }

// Focal points of Featurizer.parse_focus (/root/reference/nearl/featurizer.py:357-372): the centre of geometry of the atoms
// `sel` in each of `frames`, with numpy's arithmetic for np.mean(frame[sel], axis=0) -- the selected rows are added one after
// the other in the array's own precision, the sum is divided by the count in double and rounded to the array's precision --
// and stored as float32 like FOCALPOINTS.  Host code: the trajectory is still in host memory when the focus boxes are laid out.
int nrl_focal_means(const void *xyz, int xyz_is_f64, int n_frames_total, int n_atoms, const int *frames, int n_frames, const int *sel,
                    int n_sel, float *out) {
  if (!xyz || !frames || !sel || !out) return fail(NRL_ERR_ARGUMENT, "xyz / frames / sel / out is NULL");
  if (n_sel <= 0 || n_frames < 0 || n_atoms <= 0) return fail(NRL_ERR_SHAPE, "empty selection or trajectory");
  for (int f = 0; f < n_frames; ++f)
    if (frames[f] < 0 || frames[f] >= n_frames_total) return fail(NRL_ERR_SHAPE, "frame index outside the trajectory");
  for (int j = 0; j < n_sel; ++j)
    if (sel[j] < 0 || sel[j] >= n_atoms) return fail(NRL_ERR_SHAPE, "atom index outside the topology");
  auto impl = [&](auto *base) {
    using T = std::remove_cv_t<std::remove_pointer_t<decltype(base)>>;
    for (int f = 0; f < n_frames; ++f) {
      const T *frame = base + (size_t)frames[f] * (size_t)n_atoms * 3;
      T ax = 0, ay = 0, az = 0;
      for (int j = 0; j < n_sel; ++j) {
        const T *r = frame + (size_t)sel[j] * 3;
        ax += r[0]; ay += r[1]; az += r[2];
      }
      const double n = (double)n_sel;
      out[3 * f + 0] = (float)(T)((double)ax / n);
      out[3 * f + 1] = (float)(T)((double)ay / n);
      out[3 * f + 2] = (float)(T)((double)az / n);
    }
  };
  if (xyz_is_f64) impl(reinterpret_cast<const double *>(xyz));
  else impl(reinterpret_cast<const float *>(xyz));
  return NRL_OK;
}

// PDB text -> coordinates of every ATOM / HETATM record and the fixed-width columns of the first model's records (what
// pt.load hands /root/reference/nearl/io/traj.py:97-106 for a PDB file; nearl_b200.io.load_pdb words the semantics and is the
// fallback).  Plain host code, two passes by the caller: counts first (xyz == NULL), then the fill.  Anything that is not
// plain -- bytes outside ASCII, a bare carriage return, a record shorter than its coordinate columns, a coordinate that is not
// [blank][sign]digits[.digits][blank], models of different sizes -- returns NRL_ERR_PARSE and leaves the file to the caller.
//   columns: 16 bytes per record of the first model: atom name (cols 13-16), residue name (18-21), chain (22), residue number
//   and insertion code (23-27), element (77-78; blanks when the line ends before them).
static bool pdb_field(const char *p, double *out) {
  int i = 0;
  while (i < 8 && p[i] == ' ') ++i;
  bool neg = false;
  if (i < 8 && (p[i] == '+' || p[i] == '-')) { neg = p[i] == '-'; ++i; }
  unsigned long long mant = 0;
  int digits = 0, frac = 0;
  bool point = false;
  for (; i < 8; ++i) {
    const char ch = p[i];
    if (ch >= '0' && ch <= '9') { mant = mant * 10 + (unsigned)(ch - '0'); ++digits; if (point) ++frac; }
    else if (ch == '.' && !point) point = true;
    else break;
  }
  while (i < 8 && p[i] == ' ') ++i;
  if (i != 8 || digits == 0) return false;
  // at most 8 digits and 10^frac <= 10^7: both exact in double, one correctly rounded division = what strtod / float() give
  static const double pow10[9] = {1.0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8};
  const double v = (double)mant / pow10[frac];
  *out = neg ? -v : v;
  return true;
}

int nrl_pdb_parse(const char *text, long long n_bytes, long long *n_records, long long *n_first, long long *n_frames, double *xyz,
                  char *columns) {
  if (!text || !n_records || !n_first || !n_frames || n_bytes < 0) return fail(NRL_ERR_ARGUMENT, "text / counters is NULL");
  long long records = 0, first = 0, frames = 0, in_frame = 0;
  bool first_model = true;
  long long pos = 0;
  while (pos < n_bytes) {
    const char *line = text + pos;
    const char *nl = static_cast<const char *>(memchr(line, '\n', (size_t)(n_bytes - pos)));
    const bool has_nl = nl != nullptr;
    long long len = has_nl ? (long long)(nl - line) : n_bytes - pos;
    pos += len + (has_nl ? 1 : 0);
    if (len > 0 && line[len - 1] == '\r') --len;               // CRLF (text mode hands Python the line without it)
    for (long long k = 0; k < len; ++k)
      if ((unsigned char)line[k] >= 0x80 || line[k] == '\r') return fail(NRL_ERR_PARSE, "PDB text is not plain ASCII lines");
    if (len >= 6 && (memcmp(line, "ATOM  ", 6) == 0 || memcmp(line, "HETATM", 6) == 0)) {
      if (len < 54) return fail(NRL_ERR_PARSE, "PDB record shorter than its coordinate columns");
      double v[3];
      for (int a = 0; a < 3; ++a)
        if (!pdb_field(line + 30 + 8 * a, &v[a])) return fail(NRL_ERR_PARSE, "PDB coordinate field is not a plain decimal");
      if (xyz) { xyz[3 * records] = v[0]; xyz[3 * records + 1] = v[1]; xyz[3 * records + 2] = v[2]; }
      if (first_model) {
        if (columns) {
          char *c = columns + 16 * first;
          memcpy(c, line + 12, 4);
          memcpy(c + 4, line + 17, 4);
          memcpy(c + 8, line + 21, 6);
          // Python reads line[76:78] when len(line) >= 78, the newline counted: a column beyond the text is blank
          const bool has_elem = len + (has_nl ? 1 : 0) >= 78;
          c[14] = (has_elem && len > 76) ? line[76] : ' ';
          c[15] = (has_elem && len > 77) ? line[77] : ' ';
        }
        ++first;
      }
      ++records;
      ++in_frame;
    } else if (len >= 3 && memcmp(line, "END", 3) == 0 && in_frame > 0) {
      if (in_frame != first) return fail(NRL_ERR_PARSE, "PDB models do not list the same number of atoms");
      ++frames;
      in_frame = 0;
      first_model = false;
    }
  }
  if (in_frame > 0) {
    if (in_frame != first) return fail(NRL_ERR_PARSE, "PDB models do not list the same number of atoms");
    ++frames;
  }
  if (frames == 0) return fail(NRL_ERR_PARSE, "no ATOM / HETATM record");
  *n_records = records;
  *n_first = first;
  *n_frames = frames;
  return NRL_OK;
}

// `stream` waits (on the device) for everything enqueued on `lane` so far: a caller that spreads batches over the lanes
// joins them into its own stream without a host synchronisation (bench.py times the device-resident arm this way).
int nrl_lane_join(int lane, void *stream) {
  Ctx *c = nullptr;
  NRL_TRY(get_ctx(&c, lane));
  std::lock_guard<std::mutex> lk(c->mu);
  if (!c->join_event) NRL_CUDA(cudaEventCreateWithFlags(&c->join_event, cudaEventDisableTiming));
  NRL_CUDA(cudaEventRecord(c->join_event, c->work_stream()));
  NRL_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, c->join_event, 0));
  return NRL_OK;
}
// everything enqueued on `lane` from now on starts after what `stream` holds at this point
int nrl_lane_fork(int lane, void *stream) {
  Ctx *c = nullptr;
  NRL_TRY(get_ctx(&c, lane));
  std::lock_guard<std::mutex> lk(c->mu);
  if (!c->join_event) NRL_CUDA(cudaEventCreateWithFlags(&c->join_event, cudaEventDisableTiming));
  NRL_CUDA(cudaEventRecord(c->join_event, (cudaStream_t)stream));
  NRL_CUDA(cudaStreamWaitEvent(c->work_stream(), c->join_event, 0));
  return NRL_OK;
}

int nrl_buff_dim(float cutoff, float spacing, int *out) {
  if (!out) return fail(NRL_ERR_ARGUMENT, "out is NULL");
  Ctx *cp = nullptr;
  NRL_TRY(get_ctx(&cp));
  std::lock_guard<std::mutex> lk(cp->mu);
  NRL_TRY(cp->misc.reserve(64));
  nrl_buff_dim_kernel<<<1, 32, 0, cp->stream>>>(cutoff, spacing, cp->misc.as<int>());
  g_launches.fetch_add(1);
  NRL_CUDA(cudaGetLastError());
  NRL_CUDA(cudaMemcpyAsync(out, cp->misc.p, sizeof(int), cudaMemcpyDeviceToHost, cp->stream));
  NRL_CUDA(cudaStreamSynchronize(cp->stream));
  return NRL_OK;
}

int nrl_frame_voxelize(float *coords, const float *weights, int n_atoms, const int *dims, float spacing, float cutoff,
                       float sigma, int auto_translate, float *out) {
  if (auto_translate && coords && dims && n_atoms > 0) {
    // translate_coord, /root/reference/src/cpuutils.h:340-357 (host code in the reference too; mutates the input)
    float cog[3] = {0.f, 0.f, 0.f};
    for (int i = 0; i < n_atoms; ++i) { cog[0] += coords[i * 3]; cog[1] += coords[i * 3 + 1]; cog[2] += coords[i * 3 + 2]; }
    cog[0] /= n_atoms; cog[1] /= n_atoms; cog[2] /= n_atoms;
    const double sp = (double)spacing;
    for (int i = 0; i < n_atoms; ++i) {
      coords[i * 3] = (float)((coords[i * 3] - cog[0]) + (dims[0] * sp) / 2);
      coords[i * 3 + 1] = (float)((coords[i * 3 + 1] - cog[1]) + (dims[1] * sp) / 2);
      coords[i * 3 + 2] = (float)((coords[i * 3 + 2] - cog[2]) + (dims[2] * sp) / 2);
    }
  }
  LegacyArgs a{coords, weights, 1, n_atoms, dims, spacing, cutoff, sigma, NRL_MODE_VOXELIZE, 0, out, false, nullptr};
  return legacy_run(a);
}
int nrl_frame_voxelize_dev(const float *coords, const float *weights, int n_atoms, const int *dims, float spacing,
                           float cutoff, float sigma, float *out, void *stream) {
  LegacyArgs a{coords, weights, 1, n_atoms, dims, spacing, cutoff, sigma, NRL_MODE_VOXELIZE, 0, out, true, (cudaStream_t)stream};
  return legacy_run(a);
}

int nrl_frame_observation(const float *coords, const float *weights, int n_atoms, const int *dims, float spacing,
                          float cutoff, int type_obs, float *out) {
  if (!valid_obs(type_obs)) {   // the reference does not validate here: unknown ids give zeros
    long long G = 0;
    NRL_TRY(check_dims(dims, &G));
    if (!out) return fail(NRL_ERR_ARGUMENT, "out is NULL");
    memset(out, 0, (size_t)G * sizeof(float));
    return NRL_OK;
  }
  LegacyArgs a{coords, weights, 1, n_atoms, dims, spacing, cutoff, 0.0f, type_obs, 0, out, false, nullptr};
  return legacy_run(a);
}
int nrl_frame_observation_dev(const float *coords, const float *weights, int n_atoms, const int *dims, float spacing,
                              float cutoff, int type_obs, float *out, void *stream) {
  if (!valid_obs(type_obs)) return fail(NRL_ERR_OBSERVABLE, "The observable type is not supported");
  LegacyArgs a{coords, weights, 1, n_atoms, dims, spacing, cutoff, 0.0f, type_obs, 0, out, true, (cudaStream_t)stream};
  return legacy_run(a);
}

int nrl_marching_observer(const float *coords, const float *weights, int n_frames, int n_atoms, const int *dims,
                          float spacing, float cutoff, int type_obs, int type_agg, float *out) {
  if (!valid_obs(type_obs)) return fail(NRL_ERR_OBSERVABLE, "The observable type is not supported");
  if (!valid_agg(type_agg)) return fail(NRL_ERR_AGGREGATION, "The aggregation type is not supported");
  if (n_frames > NRL_MAX_FRAMES)
    return fail(NRL_ERR_FRAMES, "The number of frames " + std::to_string(n_frames) +
                                    " exceeds the maximum number of frames allowed " + std::to_string(NRL_MAX_FRAMES) +
                                    " frames.");
  LegacyArgs a{coords, weights, n_frames, n_atoms, dims, spacing, cutoff, 0.0f, type_obs, type_agg, out, false, nullptr};
  return legacy_run(a);
}
int nrl_marching_observer_dev(const float *coords, const float *weights, int n_frames, int n_atoms, const int *dims,
                              float spacing, float cutoff, int type_obs, int type_agg, float *out, void *stream) {
  if (!valid_obs(type_obs)) return fail(NRL_ERR_OBSERVABLE, "The observable type is not supported");
  if (!valid_agg(type_agg)) return fail(NRL_ERR_AGGREGATION, "The aggregation type is not supported");
  if (n_frames > NRL_MAX_FRAMES) return fail(NRL_ERR_FRAMES, "The number of frames exceeds the maximum number of frames allowed 512 frames.");
  LegacyArgs a{coords, weights, n_frames, n_atoms, dims, spacing, cutoff, 0.0f, type_obs, type_agg, out, true, (cudaStream_t)stream};
  return legacy_run(a);
}

int nrl_density_flow(const float *coords, const float *weights, int n_frames, int n_atoms, const int *dims,
                     float spacing, float cutoff, float sigma, int type_agg, float *out) {
  if (!valid_agg(type_agg)) return fail(NRL_ERR_AGGREGATION, "The aggregation type is not supported");
  LegacyArgs a{coords, weights, n_frames, n_atoms, dims, spacing, cutoff, sigma, NRL_MODE_VOXELIZE, type_agg, out, false, nullptr};
  return legacy_run(a);
}
int nrl_density_flow_dev(const float *coords, const float *weights, int n_frames, int n_atoms, const int *dims,
                         float spacing, float cutoff, float sigma, int type_agg, float *out, void *stream) {
  if (!valid_agg(type_agg)) return fail(NRL_ERR_AGGREGATION, "The aggregation type is not supported");
  LegacyArgs a{coords, weights, n_frames, n_atoms, dims, spacing, cutoff, sigma, NRL_MODE_VOXELIZE, type_agg, out, true, (cudaStream_t)stream};
  return legacy_run(a);
}

int nrl_aggregate(const float *series, int n_frames, long long n_voxels, int type_agg, float *out) {
  return aggregate_impl(series, n_frames, n_voxels, type_agg, out, false, nullptr);
}
int nrl_aggregate_dev(const float *series, int n_frames, long long n_voxels, int type_agg, float *out, void *stream) {
  return aggregate_impl(series, n_frames, n_voxels, type_agg, out, true, (cudaStream_t)stream);
}

int nrl_summation(const float *arr, long long n, float *out) {
  if (!out) return fail(NRL_ERR_ARGUMENT, "out is NULL");
  if (n < 0 || (n > 0 && !arr)) return fail(NRL_ERR_ARGUMENT, "arr is NULL");
  if (n == 0) { *out = 0.0f; return NRL_OK; }
  Ctx *cp = nullptr;
  NRL_TRY(get_ctx(&cp));
  Ctx &c = *cp;
  std::lock_guard<std::mutex> lk(c.mu);
  WorkspaceTurn turn(c, c.stream);
  const long long nblocks = (n + 255) / 256;
  if (nblocks >= (1ll << 31)) return fail(NRL_ERR_SHAPE, "array too long");
  NRL_TRY(c.series.reserve((size_t)n * sizeof(float)));
  NRL_TRY(c.misc.reserve((size_t)(nblocks + 16) * sizeof(float)));
  NRL_CUDA(cudaMemcpyAsync(c.series.p, arr, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, c.stream));
  float *partial = c.misc.as<float>() + 16;
  nrl_block_sum_kernel<<<(unsigned)nblocks, 256, 0, c.stream>>>(c.series.as<float>(), partial, n);
  nrl_final_sum_kernel<<<1, 32, 0, c.stream>>>(partial, (int)nblocks, c.misc.as<float>());
  g_launches.fetch_add(2);
  NRL_CUDA(cudaGetLastError());
  NRL_CUDA(cudaMemcpyAsync(out, c.misc.p, sizeof(float), cudaMemcpyDeviceToHost, c.stream));
  NRL_CUDA(cudaStreamSynchronize(c.stream));
  return NRL_OK;
}

#include "batch_api.inc"

}  // extern "C"
