// C ABI of libfast-dnn.so (include/fdnn.h): model upload, contexts, the kernel sequence of one
// forward pass, host↔device streaming.  Everything that computes runs on the GPU; without a usable
// device the compute entry points fail with FDNN_ENOGPU — there is no CPU path.
//
// One forward pass over m frames mirrors CalculationContext::Calculate (reference
// src/cpp/dnn.cc:162-165, 402-454) as a sequence of kernels on one stream:
//   input_layer → qlayer(hidden) × (L−2) → qlayer(logits) → softmax
// with u8 activations ping-ponging between two device buffers and each layer's saturation
// corrections travelling through its correction channel (device_common.cuh).

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <emmintrin.h>

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <functional>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/fdnn.h"
#include "fdnn_internal.h"
#include "kernels.h"

using namespace fdnn;

namespace {

std::atomic<long long> g_launches{0};

#define CUDA_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t e_ = (expr);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      set_error(std::string(#expr) + ": " + cudaGetErrorString(e_));                        \
      return FDNN_ECUDA;                                                                    \
    }                                                                                       \
  } while (0)

// Nothing may propagate out of an extern "C" entry point: the ones that allocate are function-try-blocks ending in this.
#define FDNN_CATCH                                                           \
  catch (const std::bad_alloc &) {                                           \
    set_error("out of host memory");                                         \
    return FDNN_ENOMEM;                                                      \
  }                                                                          \
  catch (const std::exception &e_) {                                         \
    set_error(std::string("internal error: ") + e_.what());                  \
    return FDNN_ECUDA;                                                       \
  }

int round_up(int v, int m) { return (v + m - 1) / m * m; }

// rows of logits per output-layer/softmax sub-chunk on long batches (FDNN_OUTPUT_SUB_ROWS; 0 = no sub-chunks)
const int kOutputSubRows = [] {
  const char *e = std::getenv("FDNN_OUTPUT_SUB_ROWS");
  const int v = e ? std::atoi(e) : 0;  // measured on B200 (profiles/r2_experiments.md): sub-chunks lose; off unless asked for
  return v <= 0 ? (1 << 30) : std::max(256, v / 256 * 256);
}();

bool env_flag(const char *name, bool dflt) {
  const char *e = std::getenv(name);
  if (!e || !e[0]) return dflt;
  return e[0] != '0';
}

using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// 2D u8/s8 matrix [rows][K] row-major → TMA map with a (128-byte × box_rows) box, 128B swizzle.
int make_tmap(CUtensorMap *map, const void *base, int rows, int K, int box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return FDNN_ECUDA;
  }
  cuuint64_t dims[2] = {cuuint64_t(K), cuuint64_t(rows)};
  cuuint64_t strides[1] = {cuuint64_t(K)};
  cuuint32_t box[2] = {128u, cuuint32_t(box_rows)};
  cuuint32_t estr[2] = {1u, 1u};
  static const int promo = [] {  // tuning experiments: FDNN_L2PROMO = 0 (none) / 1 (64 B) / 2 (128 B) / 3 (256 B, default)
    const char *e = std::getenv("FDNN_L2PROMO");
    return (e && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : 3;
  }();
  const CUtensorMapL2promotion promos[4] = {CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, promos[promo], CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(int(r)));
    return FDNN_ECUDA;
  }
  return FDNN_OK;
}

// The same matrix as a 3-D tensor: (128 bytes of K, row, 128-byte K block).  One box = `box_rows` rows × `box_kb` consecutive K blocks,
// delivered as box_kb consecutive 128B-swizzled tiles (the fused kernel's two-K-block pipeline stages, csrc/qlayer_fused.cu).
int make_tmap3(CUtensorMap *map, const void *base, int rows, int K, int box_rows, int box_kb) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return FDNN_ECUDA;
  }
  cuuint64_t dims[3] = {128u, cuuint64_t(rows), cuuint64_t(K / 128)};
  cuuint64_t strides[2] = {cuuint64_t(K), 128u};
  cuuint32_t box[3] = {128u, cuuint32_t(box_rows), cuuint32_t(box_kb)};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (3-D) failed with CUresult " + std::to_string(int(r)));
    return FDNN_ECUDA;
  }
  return FDNN_OK;
}

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

int usable_device(int device, int *out) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    cudaGetLastError();
    set_error("no usable CUDA device (this library has no CPU path)");
    return FDNN_ENOGPU;
  }
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) device = 0;
  }
  if (device >= count) {
    set_error("CUDA device " + std::to_string(device) + " does not exist");
    return FDNN_EINVAL;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
    cudaGetLastError();
    set_error("CUDA device " + std::to_string(device) + " is not an sm_100 (Blackwell B200) part; the kernels are built for sm_100a only");
    return FDNN_ENOGPU;
  }
  *out = device;
  return FDNN_OK;
}

}  // namespace

namespace fdnn {
bool pdl_enabled() {
  static const bool on = env_flag("FDNN_PDL", true);
  return on;
}
}  // namespace fdnn

// Inspection aid (tests/test_gpu_parity.py soak): position-weighted 64-bit checksum of a byte buffer, computed where the
// bytes are.  Every weight is odd, so a change of any single byte changes the sum; the sum is order-independent (mod 2^64),
// hence deterministic under atomics.
__global__ void digest_kernel(const uint8_t *__restrict__ data, size_t n, unsigned long long *out) {
  unsigned long long acc = 0;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
    acc += (unsigned long long) (data[i]) * ((i * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull) | 1ull);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// ---- handles ----------------------------------------------------------------------------------------

struct fdnn_model {
  int device = 0;
  int num_sms = 0;
  uint8_t *d_blob = nullptr;
  size_t blob_size = 0;
  BlobHeader hdr{};
  std::vector<BlobQLayer> q;
  std::vector<std::array<CUtensorMap, 4>> wmaps;  // per int8 layer, box rows 64 / 128 / 256 / 32
  std::vector<std::array<CUtensorMap, 3>> wmaps3;  // … as 3-D maps with boxes of 64 / 128 / 256 rows × 2 K blocks (fused kernel)
  bool fused_ok = false;                           // every int8 layer has such maps (K a multiple of 256)
  std::vector<bool> tc_ok;
  std::vector<int> fast_tail;  // per int8 layer: the packed-f32x2 epilogue is provably bit-identical (device_common.cuh)
  // certified tensor-core input layer (input_tc.cu): fixed-point limb planes and per-node statistics of layer 0
  bool input_tc = false;
  uint8_t *d_w0_limbs = nullptr;        // 3 planes of [w_plane_rows][kInputTcPitch]
  InputNodeStats *d_node_stats = nullptr;
  int w_plane_rows = 0;
  CUtensorMap w0map;
  bool force_simt = false;
  std::atomic<int> tile_policy{FDNN_POLICY_LATENCY};
  // Lifetime: the handle holds one reference and every live context one more, so a context may outlive fdnn_free of its
  // model (the reference's `delete context` never touches the dnn, jni_dnn.cc:119-133); the last release frees the device memory.
  std::atomic<int> refs{1};
  // Device group (fdnn_load_devices / FDNN_DEVICES): the primary owns one replica per device, group[0] == the primary
  // itself; fdnn_calculate shards its frames over them, contexts are handed out round-robin.  Empty for a single device.
  std::vector<fdnn_model *> group;
  std::atomic<unsigned> next_ctx{0};
  // Graph capture must not overlap device-wide synchronising calls (cudaFree, blocking copies) made
  // by this library from other threads: both sides take this lock.
  std::mutex cuda_mu;
  // contexts cached for fdnn_calculate
  std::mutex pool_mu;
  std::vector<fdnn_ctx *> pool;

  template <class T>
  const T *at(uint64_t off) const {
    return reinterpret_cast<const T *>(d_blob + off);
  }
};

struct fdnn_ctx {
  fdnn_model *model = nullptr;
  int cap = 0;  // frames
  float *d_in = nullptr;      // [cap][I]
  uint8_t *d_act[2] = {nullptr, nullptr};  // [cap][H]
  float *d_logits = nullptr;  // [cap][O]
  int8_t *d_masks = nullptr;  // [cap][O], allocated on first lazy use
  float *d_row = nullptr;     // [O] scratch for single-row lazy output
  float *d_lazy = nullptr;    // [cap][O] masked softmax rows, allocated on first batched lazy use
  uint32_t *d_fused = nullptr;  // grid-barrier counters of the fused multi-layer kernel (qlayer_fused.cu)
  // certified tensor-core input layer: transformed frames, their limb planes, row statistics, undecided elements
  float *d_xq = nullptr;
  uint8_t *d_xlimbs = nullptr;
  int x_plane_rows = 0;
  InputRowStats *d_rowstats = nullptr;
  uint32_t *d_unc_bits = nullptr;
  uint32_t *d_unc_t = nullptr;
  uint32_t *d_unc_count = nullptr;
  CUtensorMap xmap;
  bool input_tc = false;
  CUtensorMap amap[2][3];  // per activation buffer: TMA box of 128 / 64 / 32 rows (cluster 1 / 2 / 4 sharing the tile)
  CUtensorMap amap3[2];    // per activation buffer: 3-D box of 128 rows × 2 K blocks (fused kernel)
  bool amap_ok = false;
  cudaStream_t stream = nullptr;
  int policy = FDNN_POLICY_LATENCY;  // tile policy of the model when the context was created (part of every cached launch sequence)
  // fdnn_calculate on pageable caller memory (the JNI path: jni_dnn.cc:44,54-58 hands us JVM heap copies): page-locked staging
  // owned by the context, results come down in sub-chunks with one event each so that the copy-out overlaps the transfer
  float *h_in = nullptr;    // [cap][I]
  float *h_out = nullptr;   // [cap][O]
  std::vector<cudaEvent_t> events;  // polled events (wait_event): [0] = whole chunk done, [1 + k] = sub-chunk k has landed in h_out
  int8_t *h_mask = nullptr;  // single-row lazy path: mapped page-locked mask [O] and result row [O]
  float *h_row = nullptr;
  bool trace = false;
  uint8_t *d_trace = nullptr;  // [n_qlayers][cap][H]
  unsigned long long *d_timeline = nullptr;  // optional [n_qlayers][1024 CTAs][8] phase stamps (profiling aid)
  int last_frames = 0;
  bool have_logits = false;
  // One forward pass is 9 small kernels: replayed as a CUDA graph (captured on first use per
  // distinct (input, output, frames, softmax) combination) to keep launch overhead off the GPU.
  struct PassGraph {
    const float *d_in;
    float *d_out;
    int m;
    bool softmax;
    cudaGraphExec_t exec;  // nullptr: this combination has been seen once and ran un-captured; the next use captures
    int kernels;           // kernel launches one replay stands for
  };
  std::vector<PassGraph> graphs;
};

namespace {

void free_model_memory(fdnn_model *m) {
  DeviceGuard g(m->device);
  cudaFree(m->d_blob);
  cudaFree(m->d_w0_limbs);
  cudaFree(m->d_node_stats);
  delete m;
}

void model_release(fdnn_model *m) {
  if (m && m->refs.fetch_sub(1, std::memory_order_acq_rel) == 1) free_model_memory(m);
}

void destroy_ctx(fdnn_ctx *c) {
  if (!c) return;
  fdnn_model *m = c->model;
  {
    std::lock_guard<std::mutex> lk(m->cuda_mu);
    DeviceGuard g(m->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->d_in);
    cudaFree(c->d_act[0]);
    cudaFree(c->d_act[1]);
    cudaFree(c->d_logits);
    cudaFree(c->d_masks);
    cudaFree(c->d_row);
    cudaFree(c->d_lazy);
    cudaFree(c->d_trace);
    cudaFree(c->d_timeline);
    cudaFree(c->d_xq);
    cudaFree(c->d_xlimbs);
    cudaFree(c->d_rowstats);
    cudaFree(c->d_unc_bits);
    cudaFree(c->d_unc_t);
    cudaFree(c->d_unc_count);
    cudaFree(c->d_fused);
    if (c->h_in) cudaFreeHost(c->h_in);
    if (c->h_out) cudaFreeHost(c->h_out);
    if (c->h_mask) cudaFreeHost(c->h_mask);
    if (c->h_row) cudaFreeHost(c->h_row);
    for (cudaEvent_t e : c->events) cudaEventDestroy(e);
    for (auto &g2 : c->graphs)
      if (g2.exec) cudaGraphExecDestroy(g2.exec);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
  }
  model_release(m);
}

int create_ctx(fdnn_model *m, int n, fdnn_ctx **out) {
  if (n <= 0) {
    set_error("a context needs at least one frame");
    return FDNN_EINVAL;
  }
  const int I = m->hdr.in_dim, H = m->hdr.hidden, O = m->hdr.out_dim;
  // ≈ (4I + 2H + 4O) bytes of device memory per frame (+ masks and a second [n][O] buffer on lazy use)
  const double per_frame = 4.0 * I + 2.0 * H + 4.0 * O + 64;
  size_t free_b = 0, total_b = 0;
  DeviceGuard g(m->device);
  if (!g.ok) {
    set_error("cudaSetDevice failed");
    return FDNN_ECUDA;
  }
  if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && per_frame * n > 0.9 * double(free_b)) {
    set_error("context for " + std::to_string(n) + " frames does not fit in device memory");
    return FDNN_ENOMEM;
  }
  std::unique_ptr<fdnn_ctx, void (*)(fdnn_ctx *)> c(new fdnn_ctx, destroy_ctx);
  std::unique_lock<std::mutex> lk(m->cuda_mu);
  c->model = m;
  m->refs.fetch_add(1, std::memory_order_relaxed);
  c->cap = n;
  c->policy = m->tile_policy.load(std::memory_order_relaxed);
  CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaMalloc(&c->d_in, size_t(n) * I * 4));
  // activations are read by TMA in 128-row boxes; rows past `n` are never stored, but keep the
  // buffers whole tiles long so box reads stay inside the allocation's pages
  const size_t act_bytes = size_t(round_up(n, 128)) * H;
  CUDA_TRY(cudaMalloc(&c->d_act[0], act_bytes));
  CUDA_TRY(cudaMalloc(&c->d_act[1], act_bytes));
  CUDA_TRY(cudaMemsetAsync(c->d_act[0], 0, act_bytes, c->stream));
  CUDA_TRY(cudaMemsetAsync(c->d_act[1], 0, act_bytes, c->stream));
  CUDA_TRY(cudaMalloc(&c->d_logits, size_t(n) * O * 4));
  CUDA_TRY(cudaMalloc(&c->d_row, size_t(O) * 4));
  CUDA_TRY(cudaMalloc(&c->d_fused, kFusedSyncWords * sizeof(uint32_t)));
  CUDA_TRY(cudaMemsetAsync(c->d_fused, 0, kFusedSyncWords * sizeof(uint32_t), c->stream));
  if (H % 128 == 0 && !m->force_simt) {
    for (int b = 0; b < 2; ++b)
      for (int v = 0; v < 3; ++v)
        if (int rc = make_tmap(&c->amap[b][v], c->d_act[b], n, H, 128 >> v)) return rc;
    if (m->fused_ok)
      for (int b = 0; b < 2; ++b)
        if (int rc = make_tmap3(&c->amap3[b], c->d_act[b], n, H, 128, 2)) return rc;
    c->amap_ok = true;
  }
  if (m->input_tc) {
    c->x_plane_rows = round_up(n, 128);
    const size_t plane = size_t(c->x_plane_rows) * kInputTcPitch;
    CUDA_TRY(cudaMalloc(&c->d_xq, size_t(n) * I * 4));
    CUDA_TRY(cudaMalloc(&c->d_xlimbs, 3 * plane));
    CUDA_TRY(cudaMemsetAsync(c->d_xlimbs, 0, 3 * plane, c->stream));  // K padding and the rows past n stay zero
    CUDA_TRY(cudaMalloc(&c->d_rowstats, size_t(n) * sizeof(InputRowStats)));
    CUDA_TRY(cudaMalloc(&c->d_unc_bits, size_t(n) * size_t((H + 31) / 32) * sizeof(uint32_t)));
    CUDA_TRY(cudaMalloc(&c->d_unc_t, size_t((n + 31) / 32) * size_t(H) * sizeof(uint32_t)));
    CUDA_TRY(cudaMalloc(&c->d_unc_count, sizeof(uint32_t)));
    if (int rc = make_tmap(&c->xmap, c->d_xlimbs, 3 * c->x_plane_rows, kInputTcPitch, 128)) return rc;
    c->input_tc = true;
  }
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  lk.unlock();  // (a failing CUDA_TRY above unwinds lk before c, whose deleter takes the lock again)
  *out = c.release();
  return FDNN_OK;
}

// Enqueue one forward pass over frames [0, m) of `d_in` on `stream`.  Logits (lin + bias of the
// output layer) go to `d_logits` with row pitch O.
int enqueue_softmax(fdnn_ctx *c, const float *d_logits, const int8_t *d_masks, int rows, float *d_out, cudaStream_t stream);

// `want_softmax`: the caller will normalise d_logits in place right after; when the fused kernel takes the pass it does that
// itself and sets *softmax_done.
int enqueue_until_logits(fdnn_ctx *c, const float *d_in, int m, float *d_logits, cudaStream_t stream,
                         const std::function<void(int)> &after_stage = nullptr, bool want_softmax = false, bool *softmax_done = nullptr,
                         bool allow_fused = true) {
  fdnn_model *mod = c->model;
  const BlobHeader &h = mod->hdr;
  const int nq = h.n_qlayers;
  auto fix_of = [&](int layer, int variant) {
    FixList f;
    f.ptr = mod->at<uint32_t>(mod->q[size_t(layer)].off_fix_ptr[variant]);
    f.ent = mod->at<FixEntry>(mod->q[size_t(layer)].off_fix_ent[variant]);
    f.k_blocks = int(mod->q[size_t(layer)].k_blocks);
    f.group = kFixGroups[variant];
    return f;
  };

  InputLayerArgs ia{};
  ia.in = d_in;
  ia.shift = mod->at<float>(h.off_shift);
  ia.scale = mod->at<float>(h.off_scale);
  ia.w0 = mod->at<float>(h.off_w0);
  ia.bias0 = mod->at<float>(h.off_bias0);
  ia.lut = mod->at<uint8_t>(h.off_lut);
  ia.out_u8 = c->d_act[0];
  ia.M = m;
  ia.I = h.in_dim;
  ia.H = h.hidden;
  if (c->input_tc) {
    InputTcArgs ta{};
    ta.in = d_in;
    ta.shift = ia.shift;
    ta.scale = ia.scale;
    ta.w0 = ia.w0;
    ta.bias0 = ia.bias0;
    ta.lut = ia.lut;
    ta.node_stats = mod->d_node_stats;
    ta.xq = c->d_xq;
    ta.x_limbs = c->d_xlimbs;
    ta.x_plane = size_t(c->x_plane_rows) * kInputTcPitch;
    ta.x_plane_rows = c->x_plane_rows;
    ta.w_plane_rows = mod->w_plane_rows;
    ta.row_stats = c->d_rowstats;
    ta.unc_bits = c->d_unc_bits;
    ta.unc_words = (h.hidden + 31) / 32;
    ta.unc_t = c->d_unc_t;
    ta.unc_count = c->d_unc_count;
    ta.out_u8 = c->d_act[0];
    ta.M = m;
    ta.I = h.in_dim;
    ta.H = h.hidden;
    ta.fixup_ctas = mod->num_sms * 8;
    ta.num_sms = mod->num_sms;
    CUDA_TRY(launch_input_tc(c->xmap, mod->w0map, ta, stream));
    g_launches.fetch_add(3, std::memory_order_relaxed);
  } else {
    CUDA_TRY(launch_input_layer(ia, stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  if (after_stage) after_stage(0);
  const size_t act_bytes = size_t(m) * size_t(h.hidden);
  if (c->trace) CUDA_TRY(cudaMemcpyAsync(c->d_trace, c->d_act[0], act_bytes, cudaMemcpyDeviceToDevice, stream));

  // One persistent kernel for all int8 layers (+ softmax) when the batch is a single wave of tiles (qlayer_fused.cu)
  if (allow_fused && c->amap_ok && mod->fused_ok && !c->trace && nq >= 2 && nq <= kFusedMaxLayers) {
    int grid = 0;
    const int bnh = qlayer_fused_plan(m, h.hidden, h.out_dim, mod->num_sms, c->policy, &grid);
    if (bnh != 0) {
      FusedArgs fa{};
      fa.act[0] = c->amap3[0];
      fa.act[1] = c->amap3[1];
      uint32_t done = 0, tiles_so_far = 0;
      for (int j = 0; j < nq; ++j) {
        const BlobQLayer &ql = mod->q[size_t(j)];
        const bool logits = j == nq - 1;
        const int bn = logits ? 256 : bnh;
        const int variant = bn == 64 ? 0 : (bn == 128 ? 1 : 2);
        fa.w[j] = mod->wmaps3[size_t(j)][size_t(variant)];
        FusedLayer &L = fa.layer[j];
        L.bias = mod->at<float>(ql.off_bias);
        L.fix_ptr = mod->at<uint32_t>(ql.off_fix_ptr[variant]);
        L.fix_ent = mod->at<FixEntry>(ql.off_fix_ent[variant]);
        L.coeff = ql.coeff;
        L.rcp = ql.rcp_coeff;
        L.fast_div = int(ql.fast_div);
        L.fast_tail = mod->fast_tail[size_t(j)];
        L.N = ql.nodes;
        L.K = ql.inputs;
        L.need = done;
        L.n_blocks = uint32_t((ql.nodes + bn - 1) / bn);
        L.tile_begin = tiles_so_far;
        tiles_so_far += L.n_blocks * uint32_t((m + 127) / 128);
        done += L.n_blocks;
      }
      fa.n_layers = nq;
      fa.M = m;
      fa.act_buf[0] = c->d_act[0];
      fa.act_buf[1] = c->d_act[1];
      fa.out = d_logits;
      fa.out_ld = h.out_dim;
      fa.do_softmax = (want_softmax && h.out_dim <= qlayer_fused_max_softmax_width() && qlayer_fused_softmax_pays(m, grid)) ? 1 : 0;
      fa.lut = mod->at<uint8_t>(h.off_lut);
      fa.one = 1.0f;
      fa.neg_zero = -0.0f;
      fa.sync = c->d_fused;
      fa.tiles_per_row_block = done;
      fa.total_tiles = tiles_so_far;
      {
        const char *e = std::getenv("FDNN_FUSED_DEBUG");
        fa.debug_flags = e ? std::atoi(e) : 0;
      }
      fa.timeline = c->d_timeline;
      CUDA_TRY(launch_qlayer_fused(fa, bnh, grid, stream));
      g_launches.fetch_add(1, std::memory_order_relaxed);
      if (after_stage)
        for (int j = 0; j < nq; ++j) after_stage(j + 1);
      if (softmax_done) *softmax_done = fa.do_softmax != 0;
      c->last_frames = m;
      return FDNN_OK;
    }
  }

  for (int j = 0; j < nq; ++j) {
    const BlobQLayer &ql = mod->q[size_t(j)];
    const bool logits = j == nq - 1;
    QLayerArgs a{};
    a.act = c->d_act[j & 1];
    a.w = mod->at<int8_t>(ql.off_w);
    a.bias = mod->at<float>(ql.off_bias);
    a.lut = mod->at<uint8_t>(h.off_lut);
    a.coeff = ql.coeff;
    a.rcp = ql.rcp_coeff;
    a.fast_div = int(ql.fast_div);
    a.fast_tail = mod->fast_tail[size_t(j)];
    a.one = 1.0f;
    a.neg_zero = -0.0f;
    a.M = m;
    a.N = ql.nodes;
    a.K = ql.inputs;
    a.fix = fix_of(j, 0);
    {
      static const int dbg = [] {
        const char *e = std::getenv("FDNN_DEBUG");
        return e ? std::atoi(e) : 0;
      }();
      a.debug_flags = dbg;
    }
    a.timeline = c->d_timeline ? c->d_timeline + size_t(j) * 1024 * 8 : nullptr;
    if (logits) {
      a.out_f32 = d_logits;
      a.out_ld = ql.nodes;
    } else {
      a.out_u8 = c->d_act[(j + 1) & 1];
    }
    if (mod->tc_ok[size_t(j)] && c->amap_ok) {
      const TcPlan plan = qlayer_tc_plan(m, ql.nodes, logits, mod->num_sms, c->policy);
      const int which = plan.block_n == 64 ? 0 : (plan.block_n == 128 ? 1 : 2);
      a.fix = fix_of(j, which);  // the risk list grouped by the tile width
      const int act_box = plan.share_a ? (plan.cluster == 4 ? 2 : (plan.cluster == 2 ? 1 : 0)) : 0;
      const int w_rows = plan.share_a ? plan.block_n : plan.block_n / plan.cluster;
      const int w_box = w_rows == 64 ? 0 : (w_rows == 128 ? 1 : (w_rows == 256 ? 2 : 3));
      if (plan.pair && logits && want_softmax && softmax_done != nullptr && !after_stage && m > kOutputSubRows) {
        // Long batches: output layer and softmax in sub-chunks of rows whose logits (32 KB per frame at 8000 outputs) still sit in the
        // L2 when the softmax reads and overwrites them — one HBM write of the scores instead of write + read + write.  Sub-chunks
        // are whole waves of 256×256 pair tiles (9 row pairs × 32 column blocks ≈ 3.9 waves of 74 pairs for the headline network).
        const int m_pairs = (m + 255) / 256;
        const int per = std::max(1, kOutputSubRows / 256);
        const int n_sub = std::max(1, (m_pairs + per / 2) / per);
        int done_pairs = 0;
        for (int sc = 0; sc < n_sub; ++sc) {
          const int take = (m_pairs - done_pairs) / (n_sub - sc);
          QLayerArgs as = a;
          as.row0 = done_pairs * 256;
          as.M = std::min(m, (done_pairs + take) * 256);
          CUDA_TRY(launch_qlayer_pair(c->amap[j & 1][0], mod->wmaps[size_t(j)][size_t(w_box)], as, true, plan.block_n, mod->num_sms, stream));
          g_launches.fetch_add(1, std::memory_order_relaxed);
          if (int rc = enqueue_softmax(c, d_logits + size_t(as.row0) * size_t(ql.nodes), nullptr, as.M - as.row0, d_logits + size_t(as.row0) * size_t(ql.nodes), stream)) return rc;
          done_pairs += take;
        }
        *softmax_done = true;
        c->last_frames = m;
        return FDNN_OK;
      }
      if (plan.pair)
        CUDA_TRY(launch_qlayer_pair(c->amap[j & 1][0], mod->wmaps[size_t(j)][size_t(w_box)], a, logits, plan.block_n, mod->num_sms, stream));
      else
        CUDA_TRY(launch_qlayer_tc(c->amap[j & 1][act_box], mod->wmaps[size_t(j)][size_t(w_box)], a, logits, plan, mod->num_sms, stream));
    } else {
      CUDA_TRY(launch_qlayer_simt(a, logits, stream));
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (after_stage) after_stage(j + 1);
    if (c->trace && !logits)
      CUDA_TRY(cudaMemcpyAsync(c->d_trace + size_t(j + 1) * size_t(c->cap) * size_t(h.hidden), c->d_act[(j + 1) & 1], act_bytes,
                               cudaMemcpyDeviceToDevice, stream));
  }
  c->last_frames = m;
  return FDNN_OK;
}

int enqueue_softmax(fdnn_ctx *c, const float *d_logits, const int8_t *d_masks, int rows, float *d_out, cudaStream_t stream) {
  const int O = c->model->hdr.out_dim;
  SoftmaxArgs s{};
  s.logits = d_logits;
  s.mask = d_masks;
  s.out = d_out;
  s.rows = rows;
  s.O = O;
  s.ld = O;
  s.mask_ld = O;
  s.out_ld = O;
  CUDA_TRY(launch_softmax(s, stream));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return FDNN_OK;
}

// One whole pass (input layer … output layer [→ softmax in place]) on `stream`, replayed from a
// cached CUDA graph when possible.
int enqueue_pass(fdnn_ctx *c, const float *d_in, int m, float *d_out, bool softmax, cudaStream_t stream) {
  bool done = false;
  if (int rc = enqueue_until_logits(c, d_in, m, d_out, stream, nullptr, softmax, &done)) return rc;
  return (softmax && !done) ? enqueue_softmax(c, d_out, nullptr, m, d_out, stream) : FDNN_OK;
}

int run_pass(fdnn_ctx *c, const float *d_in, int m, float *d_out, bool softmax, cudaStream_t stream) {
  static const bool use_graphs = env_flag("FDNN_GRAPHS", true);
  if (!use_graphs || c->trace || c->d_timeline != nullptr || m <= 0) return enqueue_pass(c, d_in, m, d_out, softmax, stream);
  // A launch sequence is captured the SECOND time a (input, output, frames, softmax) combination shows up: callers with
  // variable-length utterances (every fdnn_calculate with a new frame count) then pay direct launches once instead of a
  // capture + instantiate per call; steady-state callers replay from their second call on.  LRU over 64 entries.
  size_t hit = c->graphs.size();
  for (size_t i = 0; i < c->graphs.size(); ++i) {
    const auto &g = c->graphs[i];
    if (g.d_in == d_in && g.d_out == d_out && g.m == m && g.softmax == softmax) {
      hit = i;
      break;
    }
  }
  if (hit < c->graphs.size() && c->graphs[hit].exec != nullptr) {
    const fdnn_ctx::PassGraph g = c->graphs[hit];
    if (hit + 1 != c->graphs.size()) {  // most recently used last
      c->graphs.erase(c->graphs.begin() + long(hit));
      c->graphs.push_back(g);
    }
    CUDA_TRY(cudaGraphLaunch(g.exec, stream));
    g_launches.fetch_add(g.kernels, std::memory_order_relaxed);
    c->last_frames = m;
    return FDNN_OK;
  }
  if (hit == c->graphs.size()) {  // first sighting: remember it, run un-captured
    if (c->graphs.size() >= 64) {
      if (c->graphs.front().exec) cudaGraphExecDestroy(c->graphs.front().exec);
      c->graphs.erase(c->graphs.begin());
    }
    c->graphs.push_back({d_in, d_out, m, softmax, nullptr, 0});
    return enqueue_pass(c, d_in, m, d_out, softmax, stream);
  }
  // Capture on the context's own stream (thread-local mode: other threads keep using CUDA freely).
  // If the capture is broken by something outside our control (another library synchronising the
  // device from a different thread), this call simply runs un-captured.
  cudaGraphExec_t exec = nullptr;
  long long kernels = 0;
  {
    std::lock_guard<std::mutex> lk(c->model->cuda_mu);
    const long long before = g_launches.load(std::memory_order_relaxed);
    cudaGraph_t graph = nullptr;
    int rc = FDNN_ECUDA;
    if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      rc = enqueue_pass(c, d_in, m, d_out, softmax, c->stream);
      if (cudaStreamEndCapture(c->stream, &graph) != cudaSuccess) rc = FDNN_ECUDA;
    }
    // capturing is not launching (other threads' launches during the capture are taken off the tally too: it is a statistic)
    kernels = g_launches.exchange(before, std::memory_order_relaxed) - before;
    if (rc == FDNN_OK && graph != nullptr && cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) exec = nullptr;
    if (graph) cudaGraphDestroy(graph);
    if (rc != FDNN_OK) exec = nullptr;
    cudaGetLastError();
  }
  if (exec == nullptr) return enqueue_pass(c, d_in, m, d_out, softmax, stream);
  c->graphs[hit].exec = exec;
  c->graphs[hit].kernels = int(kernels);
  CUDA_TRY(cudaGraphLaunch(exec, stream));
  g_launches.fetch_add(kernels, std::memory_order_relaxed);
  c->last_frames = m;
  return FDNN_OK;
}

int upload_model(const uint8_t *host_view, const void *src, bool src_on_device, size_t size, int device, fdnn_model **out) {
  std::unique_ptr<fdnn_model> m(new fdnn_model);
  std::memcpy(&m->hdr, host_view, sizeof(BlobHeader));
  m->q.resize(size_t(m->hdr.n_qlayers));
  std::memcpy(m->q.data(), host_view + m->hdr.off_qlayers, sizeof(BlobQLayer) * m->q.size());
  m->device = device;
  m->blob_size = size;
  if (m->hdr.in_dim > input_layer_max_dim()) {
    set_error("input dimension " + std::to_string(m->hdr.in_dim) + " exceeds the input-layer kernel's limit of " + std::to_string(input_layer_max_dim()));
    return FDNN_EFORMAT;
  }
  const char *env = std::getenv("FDNN_FORCE_SIMT");
  m->force_simt = env && env[0] == '1';
  DeviceGuard g(device);
  if (!g.ok) {
    set_error("cudaSetDevice failed");
    return FDNN_ECUDA;
  }
  CUDA_TRY(cudaDeviceGetAttribute(&m->num_sms, cudaDevAttrMultiProcessorCount, device));
  CUDA_TRY(cudaMalloc(&m->d_blob, size));
  cudaError_t e = cudaMemcpy(m->d_blob, src, size, src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(m->d_blob);
    set_error(std::string("model upload: ") + cudaGetErrorString(e));
    return FDNN_ECUDA;
  }
  auto fail = [&](int rc) {
    cudaFree(m->d_blob);
    cudaFree(m->d_w0_limbs);
    cudaFree(m->d_node_stats);
    return rc;
  };
  if (cudaError_t ce = input_layer_configure(); ce != cudaSuccess) {
    set_error(std::string("input_layer_configure: ") + cudaGetErrorString(ce));
    return fail(FDNN_ECUDA);
  }
  if (cudaError_t ce = qlayer_tc_configure(); ce != cudaSuccess) {
    set_error(std::string("qlayer_tc_configure: ") + cudaGetErrorString(ce));
    return fail(FDNN_ECUDA);
  }
  if (cudaError_t ce = qlayer_pair_configure(); ce != cudaSuccess) {
    set_error(std::string("qlayer_pair_configure: ") + cudaGetErrorString(ce));
    return fail(FDNN_ECUDA);
  }
  if (cudaError_t ce = qlayer_fused_configure(); ce != cudaSuccess) {
    set_error(std::string("qlayer_fused_configure: ") + cudaGetErrorString(ce));
    return fail(FDNN_ECUDA);
  }
  if (cudaError_t ce = softmax_configure(); ce != cudaSuccess) {
    set_error(std::string("softmax_configure: ") + cudaGetErrorString(ce));
    return fail(FDNN_ECUDA);
  }
  // Layer 0 for the certified tensor-core path (input_tc.cu): every weight row as a block-fixed-point integer vector
  // (|W_k| ≤ 2²², three 8-bit limbs, K padded with zeros to the plane pitch) plus scale, ‖w‖₂ and Σ|W_k| per node.
  const bool allow_input_tc = env_flag("FDNN_INPUT_TC", true);  // read per load: a test loads the same network both ways
  if (allow_input_tc && !m->force_simt && input_tc_supported(m->hdr.in_dim, m->hdr.hidden)) {
    const int I0 = m->hdr.in_dim, H0 = m->hdr.hidden;
    m->w_plane_rows = round_up(H0, 64);
    const size_t plane = size_t(m->w_plane_rows) * kInputTcPitch;
    std::vector<uint8_t> limbs(3 * plane, 0);
    std::vector<InputNodeStats> stats((size_t(H0)));
    const float *w0 = reinterpret_cast<const float *>(host_view + m->hdr.off_w0);
    const float *b0 = reinterpret_cast<const float *>(host_view + m->hdr.off_bias0);
    for (int n = 0; n < H0; ++n) {
      const float *w = w0 + size_t(n) * size_t(I0);
      float mx = 0.0f;
      bool bad = !std::isfinite(b0[n]);
      double n2 = 0.0, n2c = 0.0;
      for (int k = 0; k < I0; ++k) {
        bad = bad || !std::isfinite(w[k]);
        mx = std::max(mx, std::fabs(w[k]));
        n2 += double(w[k]) * double(w[k]);
        n2c += input_round_count(k, I0) * double(w[k]) * double(w[k]);
      }
      int e = mx > 0.0f ? std::ilogb(mx) + 1 : 0;
      if (e < -40 || e > 40) bad = true;
      if (bad) e = 0;
      double sum_abs = 0.0, sum_low = 0.0;
      for (int k = 0; k < I0 && !bad; ++k) {
        const int W = int(std::lrint(std::ldexp(double(w[k]), 22 - e)));  // exact scaling, round to nearest
        // balanced base-256 digits, every one an s8: W = d0 + 256·d1 + 65536·d2 exactly (the tensor-core kernel multiplies a frame
        // limb with all three weight limbs in one instruction, so they must share their signedness)
        const int d0 = ((W + 128) & 255) - 128, r1 = (W - d0) / 256, d1 = ((r1 + 128) & 255) - 128, d2 = (r1 - d1) / 256;
        sum_abs += std::abs(W);
        sum_low += double(std::abs(d0));
        const size_t o = size_t(n) * kInputTcPitch + size_t(k);
        limbs[o] = uint8_t(int8_t(d0));
        limbs[plane + o] = uint8_t(int8_t(d1));
        limbs[2 * plane + o] = uint8_t(int8_t(d2));
      }
      // certificate constants, every bound rounded up (derivation in input_tc.cu)
      const double u = 5.9604644775390625e-8, up = 1.0 + 1e-6;
      auto ru = [](double v) { return std::nextafter(float(v), INFINITY); };
      const double sc = std::ldexp(1.0, e - 22), nw = std::sqrt(n2) * (1.0 + 1e-9);
      InputNodeStats &s = stats[size_t(n)];
      s.c = bad ? -1.0f : float(sc);
      s.bc = bad ? 0.0f : b0[n] * 100.0f;
      s.q = ru(std::sqrt(n2c) * (1.0 + 1e-9) * up);
      s.qp = ru((nw + sc * 131072.0 * std::sqrt(double(I0))) * up);
      s.e = ru(sc * (0.5 * sum_abs + 0.25 * double(I0) + 255.0 * sum_low) * up);  // + the low-limb products the kernel skips
      s.f = ru((u * std::fabs(double(s.bc)) + 2.1 * u + 1e-9) * up);
      s.pad[0] = s.pad[1] = 0.0f;
    }
    cudaError_t ue = cudaMalloc(&m->d_w0_limbs, limbs.size());
    if (ue == cudaSuccess) ue = cudaMemcpy(m->d_w0_limbs, limbs.data(), limbs.size(), cudaMemcpyHostToDevice);
    if (ue == cudaSuccess) ue = cudaMalloc(&m->d_node_stats, stats.size() * sizeof(InputNodeStats));
    if (ue == cudaSuccess) ue = cudaMemcpy(m->d_node_stats, stats.data(), stats.size() * sizeof(InputNodeStats), cudaMemcpyHostToDevice);
    if (ue != cudaSuccess) {
      set_error(std::string("layer-0 limb upload: ") + cudaGetErrorString(ue));
      return fail(FDNN_ECUDA);
    }
    if (int rc = make_tmap(&m->w0map, m->d_w0_limbs, 3 * m->w_plane_rows, kInputTcPitch, 64)) return fail(rc);
    if (cudaError_t ce = input_tc_configure(); ce != cudaSuccess) {
      set_error(std::string("input_tc_configure: ") + cudaGetErrorString(ce));
      return fail(FDNN_ECUDA);
    }
    m->input_tc = true;
  }
  m->wmaps.resize(m->q.size());
  m->tc_ok.assign(m->q.size(), false);
  // The fast epilogue drops the reference's NaN / ≥ 2³¹ handling (x86 cvttss2si "integer indefinite"), so it
  // is only enabled where neither can occur: |sum| ≤ 255·128·K, so |lin| ≤ 255·128·K / coeff (+ 1 ulp slack).
  m->fast_tail.assign(m->q.size(), 0);
  static const bool allow_fast_tail = env_flag("FDNN_FAST_TAIL", true);
  for (size_t j = 0; j < m->q.size(); ++j) {
    const BlobQLayer &ql = m->q[j];
    const float *bias = reinterpret_cast<const float *>(host_view + ql.off_bias);
    double max_bias = 0.0;
    bool finite = std::isfinite(ql.coeff) && ql.coeff > 0.0f && std::isfinite(ql.rcp_coeff);
    for (int i = 0; i < ql.nodes && finite; ++i) {
      finite = std::isfinite(bias[i]);
      max_bias = std::max(max_bias, std::fabs(double(bias[i])));
    }
    const double max_lin = 255.0 * 128.0 * double(ql.inputs) / double(ql.coeff) * 1.001;
    m->fast_tail[j] = (allow_fast_tail && finite && ql.fast_div && (max_lin + max_bias) * 200.0 < 2.0e9) ? 1 : 0;
  }
  for (size_t j = 0; j < m->q.size(); ++j) {
    const BlobQLayer &ql = m->q[j];
    const bool logits = j + 1 == m->q.size();
    if (m->force_simt || !qlayer_tc_supported(ql.nodes, ql.inputs, logits)) continue;
    const int boxes[4] = {64, 128, 256, 32};
    for (int b = 0; b < 4; ++b)
      if (int rc = make_tmap(&m->wmaps[j][size_t(b)], m->d_blob + ql.off_w, ql.nodes, ql.inputs, boxes[b])) return fail(rc);
    m->tc_ok[j] = true;
  }
  m->wmaps3.resize(m->q.size());
  m->fused_ok = !m->q.empty();
  for (size_t j = 0; j < m->q.size(); ++j) {
    const BlobQLayer &ql = m->q[j];
    if (!m->tc_ok[j] || ql.inputs % 256 != 0) {
      m->fused_ok = false;
      break;
    }
    for (int b = 0; b < 3; ++b)
      if (int rc = make_tmap3(&m->wmaps3[j][size_t(b)], m->d_blob + ql.off_w, ql.nodes, ql.inputs, 64 << b, 2)) return fail(rc);
  }
  *out = m.release();
  return FDNN_OK;
}

}  // namespace

// ---- exported C ABI ---------------------------------------------------------------------------------

extern "C" {

const char *fdnn_last_error(void) { return get_error(); }
const char *fdnn_version(void) { return "fast-dnn-b200 0.1 (sm_100a)"; }

int fdnn_pack(const char *path, float cutoff, void **blob, size_t *size) try {
  if (!blob || !size) {
    set_error("null output pointer");
    return FDNN_EINVAL;
  }
  std::vector<uint8_t> v;
  if (int rc = pack_model(path, cutoff, v)) return rc;
  void *p = std::malloc(v.size());
  if (!p) {
    set_error("out of host memory");
    return FDNN_ENOMEM;
  }
  std::memcpy(p, v.data(), v.size());
  *blob = p;
  *size = v.size();
  return FDNN_OK;
} FDNN_CATCH

void fdnn_blob_free(void *blob) { std::free(blob); }

int fdnn_align_dnn_bin(const char *in_path, const char *out_path, int input_alignment, int hidden_alignment) try {
  return align_dnn_bin(in_path, out_path, input_alignment, hidden_alignment);
} FDNN_CATCH

int fdnn_import_kaldi_nnet1(const char *nnet_txt_path, const char *transform_txt_path, const char *out_dnn_bin_path) try {
  return import_kaldi_nnet1(nnet_txt_path, transform_txt_path, out_dnn_bin_path);
} FDNN_CATCH

int fdnn_feature_bin_read(const char *path, int *frames, int *dim, float **data) try {
  if (!path || !frames || !dim || !data) {
    set_error("null argument");
    return FDNN_EINVAL;
  }
  std::vector<float> v;
  if (int rc = read_feature_bin(path, frames, dim, v)) return rc;
  float *p = static_cast<float *>(std::malloc(std::max<size_t>(v.size(), 1) * sizeof(float)));
  if (!p) {
    set_error("out of host memory");
    return FDNN_ENOMEM;
  }
  if (!v.empty()) std::memcpy(p, v.data(), v.size() * sizeof(float));
  *data = p;
  return FDNN_OK;
} FDNN_CATCH

int fdnn_feature_bin_write(const char *path, const float *data, int frames, int dim) try { return write_feature_bin(path, data, frames, dim); } FDNN_CATCH
int fdnn_output_dump_write(const char *path, const float *data, int frames, int dim) try { return write_output_dump(path, data, frames, dim); } FDNN_CATCH

int fdnn_load_blob(const void *blob, size_t size, int device, fdnn_model **out) try {
  if (!blob || !out) {
    set_error("null argument");
    return FDNN_EINVAL;
  }
  int dev = 0;
  if (int rc = usable_device(device, &dev)) return rc;
  cudaPointerAttributes attr{};
  bool on_device = cudaPointerGetAttributes(&attr, blob) == cudaSuccess && attr.type == cudaMemoryTypeDevice;
  cudaGetLastError();
  std::vector<uint8_t> host_copy;
  const uint8_t *view = static_cast<const uint8_t *>(blob);
  if (on_device) {
    // validation needs the index sections on the host; weights stay on the device
    host_copy.resize(size);
    DeviceGuard g(dev);
    CUDA_TRY(cudaMemcpy(host_copy.data(), blob, size, cudaMemcpyDeviceToHost));
    view = host_copy.data();
  }
  if (int rc = validate_blob(view, size)) return rc;
  return upload_model(view, blob, on_device, size, dev, out);
} FDNN_CATCH

// ---- device groups: one replica per GPU behind ONE handle ----------------------------------------------
// The host parses and quantizes once, the packed blob goes to the first device, ONE ncclBroadcast (in-process communicator,
// ncclCommInitAll) delivers it to the others, and every device builds its replica from its copy.  No collective afterwards:
// fdnn_calculate cuts its frames into one contiguous shard per device.  NCCL is bound at run time (dlopen "libnccl.so.2"),
// so a single-GPU JVM needs no NCCL installed; a device group without it fails loudly.
}  // extern "C"

namespace {

struct NcclApi {
  using Comm = void *;
  int (*CommInitAll)(Comm *, int, const int *) = nullptr;
  int (*CommDestroy)(Comm) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Broadcast)(const void *, void *, size_t, int, int, Comm, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool ok = false;
};

const NcclApi &nccl_api() {
  static const NcclApi api = [] {
    NcclApi a;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return a;
    a.CommInitAll = reinterpret_cast<decltype(a.CommInitAll)>(dlsym(h, "ncclCommInitAll"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(dlsym(h, "ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(dlsym(h, "ncclGroupEnd"));
    a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(dlsym(h, "ncclBroadcast"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    a.ok = a.CommInitAll && a.CommDestroy && a.GroupStart && a.GroupEnd && a.Broadcast && a.GetErrorString;
    return a;
  }();
  return api;
}

std::atomic<long long> g_nccl_broadcasts{0};

// blob (host) → one device buffer per device of `devs`; devs[0] gets it over PCIe, the rest by one ncclBroadcast
int broadcast_blob(const std::vector<uint8_t> &blob, const std::vector<int> &devs, std::vector<uint8_t *> &d_out) {
  const NcclApi &nccl = nccl_api();
  if (!nccl.ok) {
    set_error("a device group needs NCCL (libnccl.so.2 could not be loaded): the weight blob is delivered by one ncclBroadcast");
    return FDNN_ECUDA;
  }
  const int n = int(devs.size());
  d_out.assign(size_t(n), nullptr);
  std::vector<cudaStream_t> streams(size_t(n), nullptr);
  std::vector<NcclApi::Comm> comms(size_t(n), nullptr);
  int rc = FDNN_OK;
  auto cuda_ok = [&](cudaError_t e, const char *what) {
    if (e != cudaSuccess && rc == FDNN_OK) {
      set_error(std::string(what) + ": " + cudaGetErrorString(e));
      rc = FDNN_ECUDA;
    }
    return e == cudaSuccess;
  };
  auto nccl_ok = [&](int e, const char *what) {
    if (e != 0 && rc == FDNN_OK) {
      set_error(std::string(what) + ": " + nccl.GetErrorString(e));
      rc = FDNN_ECUDA;
    }
    return e == 0;
  };
  int prev = -1;
  cudaGetDevice(&prev);
  for (int i = 0; i < n && rc == FDNN_OK; ++i) {
    cuda_ok(cudaSetDevice(devs[size_t(i)]), "cudaSetDevice");
    cuda_ok(cudaMalloc(&d_out[size_t(i)], blob.size()), "cudaMalloc(blob)");
    cuda_ok(cudaStreamCreateWithFlags(&streams[size_t(i)], cudaStreamNonBlocking), "cudaStreamCreate");
  }
  if (rc == FDNN_OK) {
    cuda_ok(cudaSetDevice(devs[0]), "cudaSetDevice");
    cuda_ok(cudaMemcpyAsync(d_out[0], blob.data(), blob.size(), cudaMemcpyHostToDevice, streams[0]), "blob upload");
  }
  if (rc == FDNN_OK && nccl_ok(nccl.CommInitAll(comms.data(), n, devs.data()), "ncclCommInitAll")) {
    nccl_ok(nccl.GroupStart(), "ncclGroupStart");
    for (int i = 0; i < n && rc == FDNN_OK; ++i) {
      cuda_ok(cudaSetDevice(devs[size_t(i)]), "cudaSetDevice");
      nccl_ok(nccl.Broadcast(d_out[0], d_out[size_t(i)], blob.size(), /*ncclUint8*/ 1, /*root*/ 0, comms[size_t(i)], streams[size_t(i)]),
              "ncclBroadcast");
    }
    nccl_ok(nccl.GroupEnd(), "ncclGroupEnd");
    if (rc == FDNN_OK) g_nccl_broadcasts.fetch_add(1, std::memory_order_relaxed);
  }
  for (int i = 0; i < n; ++i) {
    if (streams[size_t(i)]) {
      cudaSetDevice(devs[size_t(i)]);
      cuda_ok(cudaStreamSynchronize(streams[size_t(i)]), "blob broadcast");
    }
  }
  for (int i = 0; i < n; ++i) {
    if (comms[size_t(i)]) nccl.CommDestroy(comms[size_t(i)]);
    if (streams[size_t(i)]) {
      cudaSetDevice(devs[size_t(i)]);
      cudaStreamDestroy(streams[size_t(i)]);
    }
  }
  if (rc != FDNN_OK)
    for (int i = 0; i < n; ++i)
      if (d_out[size_t(i)]) {
        cudaSetDevice(devs[size_t(i)]);
        cudaFree(d_out[size_t(i)]);
        d_out[size_t(i)] = nullptr;
      }
  if (prev >= 0) cudaSetDevice(prev);
  return rc;
}

// "all", "0,1,2,3", "4" (device 4 alone) → device list; empty = not a group
std::vector<int> parse_device_list(const char *text) {
  std::vector<int> devs;
  if (!text || !text[0]) return devs;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) {
    cudaGetLastError();
    return devs;
  }
  const std::string t(text);
  if (t == "all" || t == "ALL") {
    for (int i = 0; i < count; ++i) devs.push_back(i);
    return devs;
  }
  size_t pos = 0;
  while (pos < t.size()) {
    size_t end = t.find(',', pos);
    if (end == std::string::npos) end = t.size();
    const std::string item = t.substr(pos, end - pos);
    if (!item.empty()) devs.push_back(std::atoi(item.c_str()));
    pos = end + 1;
  }
  return devs;
}

void release_group(fdnn_model *model) {
  // pooled contexts first (each holds a reference on its replica), then the replicas' handle references
  std::vector<fdnn_model *> members = model->group.empty() ? std::vector<fdnn_model *>{model} : model->group;
  for (fdnn_model *r : members) {
    std::vector<fdnn_ctx *> pooled;
    {
      std::lock_guard<std::mutex> lk(r->pool_mu);
      pooled.swap(r->pool);
    }
    for (fdnn_ctx *c : pooled) destroy_ctx(c);
  }
  for (fdnn_model *r : members)
    if (r != model) model_release(r);
  model_release(model);
}

}  // namespace

extern "C" {

int fdnn_load_devices(const char *path, float cutoff, const int *devices, int n_devices, fdnn_model **out) try {
  if (!out || !devices || n_devices <= 0) {
    set_error("bad argument to fdnn_load_devices");
    return FDNN_EINVAL;
  }
  std::vector<int> devs;
  for (int i = 0; i < n_devices; ++i) {
    int dev = 0;
    if (int rc = usable_device(devices[i], &dev)) return rc;
    if (std::find(devs.begin(), devs.end(), dev) != devs.end()) {
      set_error("device " + std::to_string(dev) + " is listed twice");
      return FDNN_EINVAL;
    }
    devs.push_back(dev);
  }
  std::vector<uint8_t> v;
  if (int rc = pack_model(path, cutoff, v)) return rc;
  if (devs.size() == 1) return upload_model(v.data(), v.data(), false, v.size(), devs[0], out);
  std::vector<uint8_t *> d_blobs;
  if (int rc = broadcast_blob(v, devs, d_blobs)) return rc;
  std::vector<fdnn_model *> reps;
  int rc = FDNN_OK;
  for (size_t i = 0; i < devs.size() && rc == FDNN_OK; ++i) {
    fdnn_model *r = nullptr;
    rc = upload_model(v.data(), d_blobs[i], true, v.size(), devs[i], &r);  // index sections from the host copy, weights device → device
    if (rc == FDNN_OK) reps.push_back(r);
  }
  for (size_t i = 0; i < devs.size(); ++i) {
    DeviceGuard g(devs[i]);
    cudaFree(d_blobs[i]);
  }
  if (rc != FDNN_OK) {
    for (fdnn_model *r : reps) model_release(r);
    return rc;
  }
  reps[0]->group = reps;
  *out = reps[0];
  return FDNN_OK;
} FDNN_CATCH

int fdnn_load(const char *path, float cutoff, int device, fdnn_model **out) try {
  if (!out) {
    set_error("null output pointer");
    return FDNN_EINVAL;
  }
  // The JNI surface has no device argument (jni_dnn.cc:7-18): FDNN_DEVICES=all | 0,1,2,3 makes the handle that
  // Java_suskun_nn_QuantizedDnn_initialize returns a device group.
  if (device < 0) {
    const std::vector<int> devs = parse_device_list(std::getenv("FDNN_DEVICES"));
    if (!devs.empty()) return fdnn_load_devices(path, cutoff, devs.data(), int(devs.size()), out);
  }
  int dev = 0;
  if (int rc = usable_device(device, &dev)) return rc;
  std::vector<uint8_t> v;
  if (int rc = pack_model(path, cutoff, v)) return rc;
  return upload_model(v.data(), v.data(), false, v.size(), dev, out);
} FDNN_CATCH

int fdnn_free(fdnn_model *model) {
  if (!model) return FDNN_OK;
  release_group(model);
  return FDNN_OK;
}

int fdnn_device_count(const fdnn_model *m) { return m ? (m->group.empty() ? 1 : int(m->group.size())) : FDNN_EINVAL; }
int fdnn_device_at(const fdnn_model *m, int i) {
  if (!m || i < 0 || i >= fdnn_device_count(m)) return FDNN_EINVAL;
  return m->group.empty() ? m->device : m->group[size_t(i)]->device;
}
long long fdnn_nccl_broadcast_count(void) { return g_nccl_broadcasts.load(std::memory_order_relaxed); }

int fdnn_input_dim(const fdnn_model *m) { return m ? m->hdr.in_dim : FDNN_EINVAL; }
int fdnn_output_dim(const fdnn_model *m) { return m ? m->hdr.out_dim : FDNN_EINVAL; }
int fdnn_hidden_dim(const fdnn_model *m) { return m ? m->hdr.hidden : FDNN_EINVAL; }
int fdnn_device(const fdnn_model *m) { return m ? m->device : FDNN_EINVAL; }
int fdnn_set_tile_policy(fdnn_model *m, int policy) {
  if (!m || (policy != FDNN_POLICY_LATENCY && policy != FDNN_POLICY_THROUGHPUT)) {
    set_error("bad argument to fdnn_set_tile_policy");
    return FDNN_EINVAL;
  }
  m->tile_policy.store(policy, std::memory_order_relaxed);
  for (fdnn_model *r : m->group) r->tile_policy.store(policy, std::memory_order_relaxed);
  return FDNN_OK;
}
int fdnn_layer_count(const fdnn_model *m) { return m ? m->hdr.n_qlayers + 1 : FDNN_EINVAL; }

int fdnn_layer_dim(const fdnn_model *m, int i) {
  if (!m) return FDNN_EINVAL;
  if (i == 0) return m->hdr.hidden;
  // jni_dnn.cc:135-148 indexes the int8 layer vector with i itself: valid for 1 ≤ i < n_qlayers,
  // −1 past it (and where the reference would read out of bounds, i == n_qlayers)
  if (i < 0 || i >= m->hdr.n_qlayers) return -1;
  return m->q[size_t(i)].nodes;
}

int fdnn_model_qlayer(const fdnn_model *m, int i, int *nodes, int *inputs, float *multiplier, int8_t *weights, float *bias) try {
  if (!m || i < 0 || i >= m->hdr.n_qlayers) {
    set_error("bad layer index");
    return FDNN_EINVAL;
  }
  const BlobQLayer &q = m->q[size_t(i)];
  if (nodes) *nodes = q.nodes;
  if (inputs) *inputs = q.inputs;
  if (multiplier) *multiplier = q.multiplier;
  std::lock_guard<std::mutex> lk(const_cast<fdnn_model *>(m)->cuda_mu);
  DeviceGuard g(m->device);
  if (weights) CUDA_TRY(cudaMemcpy(weights, m->d_blob + q.off_w, size_t(q.nodes) * size_t(q.inputs), cudaMemcpyDeviceToHost));
  if (bias) CUDA_TRY(cudaMemcpy(bias, m->d_blob + q.off_bias, size_t(q.nodes) * 4, cudaMemcpyDeviceToHost));
  return FDNN_OK;
} FDNN_CATCH

int fdnn_model_fixup_count(const fdnn_model *m, int i) {
  if (!m || i < 0 || i >= m->hdr.n_qlayers) return FDNN_EINVAL;
  return int(m->q[size_t(i)].n_fix);
}

int fdnn_model_fast_div(const fdnn_model *m, int i) {
  if (!m || i < 0 || i >= m->hdr.n_qlayers) return FDNN_EINVAL;
  return int(m->q[size_t(i)].fast_div);
}

int fdnn_model_uses_tensor_cores(const fdnn_model *m, int i) {
  if (!m || i < 0 || i >= m->hdr.n_qlayers) return FDNN_EINVAL;
  return m->tc_ok[size_t(i)] ? 1 : 0;
}

int fdnn_sigmoid_lut(uint8_t out[1280]) {
  if (!out) return FDNN_EINVAL;
  build_reference_lut(out);
  return FDNN_OK;
}

// ---- contexts -----------------------------------------------------------------------------------------

int fdnn_ctx_new(fdnn_model *model, int n, int batch_hint, fdnn_ctx **out) try {
  (void) batch_hint;
  if (!model || !out) {
    set_error("null argument");
    return FDNN_EINVAL;
  }
  // a device group hands its contexts out round-robin: one context lives on one GPU
  fdnn_model *home = model->group.empty() ? model : model->group[model->next_ctx.fetch_add(1, std::memory_order_relaxed) % model->group.size()];
  return create_ctx(home, n, out);
} FDNN_CATCH

int fdnn_ctx_free(fdnn_ctx *ctx) {
  destroy_ctx(ctx);
  return FDNN_OK;
}

int fdnn_ctx_frames(const fdnn_ctx *ctx) { return ctx ? ctx->cap : FDNN_EINVAL; }
int fdnn_ctx_output_dim(const fdnn_ctx *ctx) { return ctx ? ctx->model->hdr.out_dim : FDNN_EINVAL; }
int fdnn_ctx_input_dim(const fdnn_ctx *ctx) { return ctx ? ctx->model->hdr.in_dim : FDNN_EINVAL; }

// Profiling aid: per-CTA phase timestamps (SM clocks) of the tensor-core layer kernels of the NEXT
// forward pass.  out = [n_qlayers][1024][8] uint64 (host); slots: 0 entry, 1 setup done, 2 first
// operands landed, 3 last MMA committed, 4 epilogue staging done, 5 accumulator ready, 6 tile done,
// 7 exit.  enable = 1 arms, enable = 0 copies the stamps out and disarms.
int fdnn_ctx_timeline(fdnn_ctx *ctx, int enable, unsigned long long *out) {
  if (!ctx) return FDNN_EINVAL;
  std::lock_guard<std::mutex> lk(ctx->model->cuda_mu);
  DeviceGuard g(ctx->model->device);
  const size_t bytes = size_t(ctx->model->hdr.n_qlayers) * 1024 * 8 * sizeof(unsigned long long);
  if (enable) {
    if (!ctx->d_timeline) CUDA_TRY(cudaMalloc(&ctx->d_timeline, bytes));
    CUDA_TRY(cudaMemset(ctx->d_timeline, 0, bytes));
    CUDA_TRY(cudaDeviceSynchronize());
    return FDNN_OK;
  }
  if (!ctx->d_timeline || !out) return FDNN_EINVAL;
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(out, ctx->d_timeline, bytes, cudaMemcpyDeviceToHost));
  cudaFree(ctx->d_timeline);
  ctx->d_timeline = nullptr;
  return FDNN_OK;
}

int fdnn_ctx_set_trace(fdnn_ctx *ctx, int enable) {
  if (!ctx) return FDNN_EINVAL;
  DeviceGuard g(ctx->model->device);
  if (enable && !ctx->d_trace)
    CUDA_TRY(cudaMalloc(&ctx->d_trace, size_t(ctx->model->hdr.n_qlayers) * size_t(ctx->cap) * size_t(ctx->model->hdr.hidden)));
  ctx->trace = enable != 0;
  return FDNN_OK;
}

int fdnn_ctx_until_output_device(fdnn_ctx *ctx, const float *d_in, int n_frames, void *stream) try {
  if (!ctx || !d_in || n_frames < 0 || n_frames > ctx->cap) {
    set_error("bad argument to fdnn_ctx_until_output_device");
    return FDNN_EINVAL;
  }
  DeviceGuard g(ctx->model->device);
  if (int rc = run_pass(ctx, d_in, n_frames, ctx->d_logits, false, static_cast<cudaStream_t>(stream))) return rc;
  ctx->have_logits = true;
  return FDNN_OK;
} FDNN_CATCH

int fdnn_ctx_forward_device(fdnn_ctx *ctx, const float *d_in, int n_frames, float *d_out, void *stream) try {
  if (!ctx || !d_in || !d_out || n_frames < 0 || n_frames > ctx->cap) {
    set_error("bad argument to fdnn_ctx_forward_device");
    return FDNN_EINVAL;
  }
  DeviceGuard g(ctx->model->device);
  ctx->have_logits = false;  // logits are produced straight into the caller's buffer and normalised in place
  return run_pass(ctx, d_in, n_frames, d_out, true, static_cast<cudaStream_t>(stream));
} FDNN_CATCH

int fdnn_ctx_lazy_batch_device(fdnn_ctx *ctx, const int8_t *d_masks, int n_frames, float *d_out, void *stream) {
  if (!ctx || !d_masks || !d_out || n_frames < 0 || n_frames > ctx->last_frames) {
    set_error("bad argument to fdnn_ctx_lazy_batch_device");
    return FDNN_EINVAL;
  }
  if (!ctx->have_logits) {
    set_error("calculateUntilOutput has not been run on this context");
    return FDNN_EINVAL;
  }
  DeviceGuard g(ctx->model->device);
  return enqueue_softmax(ctx, ctx->d_logits, d_masks, n_frames, d_out, static_cast<cudaStream_t>(stream));
}

int fdnn_ctx_until_output(fdnn_ctx *ctx, const float *in) try {
  if (!ctx || !in) {
    set_error("null argument");
    return FDNN_EINVAL;
  }
  DeviceGuard g(ctx->model->device);
  const size_t bytes = size_t(ctx->cap) * size_t(ctx->model->hdr.in_dim) * 4;
  CUDA_TRY(cudaMemcpyAsync(ctx->d_in, in, bytes, cudaMemcpyHostToDevice, ctx->stream));
  if (int rc = run_pass(ctx, ctx->d_in, ctx->cap, ctx->d_logits, false, ctx->stream)) return rc;
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  ctx->have_logits = true;
  return FDNN_OK;
} FDNN_CATCH

int fdnn_ctx_lazy(fdnn_ctx *ctx, int idx, const int8_t *mask, float *out) try {
  if (!ctx || !mask || !out) {
    set_error("null argument");
    return FDNN_EINVAL;
  }
  if (!ctx->have_logits) {
    set_error("calculateUntilOutput has not been run on this context");
    return FDNN_EINVAL;
  }
  if (idx < 0 || idx >= ctx->last_frames) {
    set_error("frame index out of range");
    return FDNN_EINVAL;
  }
  DeviceGuard g(ctx->model->device);
  const int O = ctx->model->hdr.out_dim;
  // One frame per call is a latency path (the Java side calls this once per frame, QuantizedDnn.java:88-93): no copy
  // engine round trips.  The mask sits in mapped page-locked memory the kernel reads over PCIe, the kernel writes the
  // row straight into mapped page-locked memory, and the only host↔device handshake is one stream synchronise.
  if (!ctx->h_mask) {
    CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&ctx->h_mask), size_t(O), cudaHostAllocMapped));
    CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&ctx->h_row), size_t(O) * 4, cudaHostAllocMapped));
  }
  std::memcpy(ctx->h_mask, mask, size_t(O));
  if (int rc = enqueue_softmax(ctx, ctx->d_logits + size_t(idx) * size_t(O), ctx->h_mask, 1, ctx->h_row, ctx->stream)) return rc;
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  std::memcpy(out, ctx->h_row, size_t(O) * 4);
  return FDNN_OK;
} FDNN_CATCH

int fdnn_ctx_hidden(fdnn_ctx *ctx, int layer, int n_frames, uint8_t *out) {
  if (!ctx || !out) return FDNN_EINVAL;
  const int nq = ctx->model->hdr.n_qlayers, H = ctx->model->hdr.hidden;
  if (layer < 0 || layer > nq - 1 || n_frames < 0 || n_frames > ctx->last_frames) {
    set_error("bad layer or frame count");
    return FDNN_EINVAL;
  }
  DeviceGuard g(ctx->model->device);
  const uint8_t *src = nullptr;
  if (layer == nq - 1)
    src = ctx->d_act[(nq - 1) & 1];
  else if (ctx->trace && ctx->d_trace)
    src = ctx->d_trace + size_t(layer) * size_t(ctx->cap) * size_t(H);
  else {
    set_error("only the last hidden layer is retained unless trace mode was enabled before the forward pass");
    return FDNN_EINVAL;
  }
  std::lock_guard<std::mutex> lk(ctx->model->cuda_mu);
  CUDA_TRY(cudaMemcpy(out, src, size_t(n_frames) * size_t(H), cudaMemcpyDeviceToHost));
  return FDNN_OK;
}

int fdnn_ctx_hidden_digest(fdnn_ctx *ctx, int layer, int n_frames, unsigned long long *digest) {
  if (!ctx || !digest) return FDNN_EINVAL;
  const int nq = ctx->model->hdr.n_qlayers, H = ctx->model->hdr.hidden;
  if (layer < 0 || layer > nq - 1 || n_frames < 0 || n_frames > ctx->last_frames) {
    set_error("bad layer or frame count");
    return FDNN_EINVAL;
  }
  const uint8_t *src = nullptr;
  if (layer == nq - 1)
    src = ctx->d_act[(nq - 1) & 1];
  else if (ctx->trace && ctx->d_trace)
    src = ctx->d_trace + size_t(layer) * size_t(ctx->cap) * size_t(H);
  else {
    set_error("only the last hidden layer is retained unless trace mode was enabled before the forward pass");
    return FDNN_EINVAL;
  }
  DeviceGuard g(ctx->model->device);
  std::lock_guard<std::mutex> lk(ctx->model->cuda_mu);
  unsigned long long *d_sum = nullptr;
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMalloc(&d_sum, sizeof(unsigned long long)));
  cudaMemset(d_sum, 0, sizeof(unsigned long long));
  digest_kernel<<<ctx->model->num_sms * 8, 256>>>(src, size_t(n_frames) * size_t(H), d_sum);
  cudaError_t e = cudaMemcpy(digest, d_sum, sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  cudaFree(d_sum);
  if (e != cudaSuccess) {
    set_error(std::string("hidden digest: ") + cudaGetErrorString(e));
    return FDNN_ECUDA;
  }
  return FDNN_OK;
}

int fdnn_ctx_logits(fdnn_ctx *ctx, int n_frames, float *out) {
  if (!ctx || !out || n_frames < 0 || n_frames > ctx->last_frames || !ctx->have_logits) {
    set_error("no resident logits for that many frames");
    return FDNN_EINVAL;
  }
  std::lock_guard<std::mutex> lk(ctx->model->cuda_mu);
  DeviceGuard g(ctx->model->device);
  CUDA_TRY(cudaMemcpy(out, ctx->d_logits, size_t(n_frames) * size_t(ctx->model->hdr.out_dim) * 4, cudaMemcpyDeviceToHost));
  return FDNN_OK;
}

// Bench/profiling aid: `iters` forward passes over device-resident input with CUDA events between
// the kernels.  ms[0] = input layer, ms[1 .. nq−1] = hidden int8 layers, ms[nq] = output int8
// layer, ms[nq+1] = softmax (averages, milliseconds).  Event pairs between back-to-back kernels
// include the inter-kernel launch gap.
int fdnn_ctx_input_undecided(fdnn_ctx *ctx, unsigned *undecided) {
  if (!ctx || !undecided) {
    set_error("bad argument to fdnn_ctx_input_undecided");
    return FDNN_EINVAL;
  }
  *undecided = 0xffffffffu;
  if (!ctx->input_tc) return FDNN_OK;
  DeviceGuard g(ctx->model->device);
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(undecided, ctx->d_unc_count, sizeof(unsigned), cudaMemcpyDeviceToHost));
  return FDNN_OK;
}

int fdnn_ctx_profile_stages(fdnn_ctx *ctx, const float *d_in, int n_frames, float *d_out, int iters, float *ms) try {
  if (!ctx || !d_in || !d_out || !ms || iters <= 0 || n_frames <= 0 || n_frames > ctx->cap) {
    set_error("bad argument to fdnn_ctx_profile_stages");
    return FDNN_EINVAL;
  }
  DeviceGuard g(ctx->model->device);
  const int stages = ctx->model->hdr.n_qlayers + 2;
  std::vector<cudaEvent_t> ev(size_t(stages) + 1);
  for (auto &e : ev) CUDA_TRY(cudaEventCreate(&e));
  std::vector<double> total(size_t(stages), 0.0);
  int rc = FDNN_OK;
  for (int it = 0; it < iters && rc == FDNN_OK; ++it) {
    cudaEventRecord(ev[0], ctx->stream);
    rc = enqueue_until_logits(ctx, d_in, n_frames, d_out, ctx->stream, [&](int stage) { cudaEventRecord(ev[size_t(stage) + 1], ctx->stream); },
                              false, nullptr, /*allow_fused=*/false);
    if (rc == FDNN_OK) rc = enqueue_softmax(ctx, d_out, nullptr, n_frames, d_out, ctx->stream);
    cudaEventRecord(ev[size_t(stages)], ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      set_error("profile_stages: stream synchronize failed");
      rc = FDNN_ECUDA;
    }
    for (int s = 0; s < stages && rc == FDNN_OK; ++s) {
      float t = 0;
      cudaEventElapsedTime(&t, ev[size_t(s)], ev[size_t(s) + 1]);
      total[size_t(s)] += t;
    }
  }
  ctx->have_logits = false;
  for (auto &e : ev) cudaEventDestroy(e);
  for (int s = 0; s < stages; ++s) ms[s] = float(total[size_t(s)] / iters);
  return rc;
} FDNN_CATCH

// Bench/profiling aid: the pass as it normally runs (fused kernel where it applies), `iters` times on the context's stream
// with CUDA events after the input layer and at the end.  ms[0] = fp32 input layer, ms[1] = everything after it (the fused
// int8 stack + softmax: ONE kernel when *fused = 1); averages in milliseconds.
int fdnn_ctx_profile_pass(fdnn_ctx *ctx, const float *d_in, int n_frames, float *d_out, int iters, float *ms, int *fused) try {
  if (!ctx || !d_in || !d_out || !ms || iters <= 0 || n_frames <= 0 || n_frames > ctx->cap) {
    set_error("bad argument to fdnn_ctx_profile_pass");
    return FDNN_EINVAL;
  }
  DeviceGuard g(ctx->model->device);
  cudaEvent_t ev[3];
  for (auto &e : ev) CUDA_TRY(cudaEventCreate(&e));
  double total[2] = {0.0, 0.0};
  int rc = FDNN_OK;
  bool took_fused = false;
  for (int it = 0; it < iters && rc == FDNN_OK; ++it) {
    const long long before = g_launches.load(std::memory_order_relaxed);
    long long after_input = before;
    bool done = false;
    cudaEventRecord(ev[0], ctx->stream);
    rc = enqueue_until_logits(ctx, d_in, n_frames, d_out, ctx->stream,
                              [&](int stage) {
                                if (stage == 0) {
                                  cudaEventRecord(ev[1], ctx->stream);
                                  after_input = g_launches.load(std::memory_order_relaxed);
                                }
                              },
                              true, &done);
    if (rc == FDNN_OK && !done) rc = enqueue_softmax(ctx, d_out, nullptr, n_frames, d_out, ctx->stream);
    cudaEventRecord(ev[2], ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      set_error("profile_pass: stream synchronize failed");
      rc = FDNN_ECUDA;
    }
    took_fused = g_launches.load(std::memory_order_relaxed) - after_input == 1;
    for (int s2 = 0; s2 < 2 && rc == FDNN_OK; ++s2) {
      float t = 0;
      cudaEventElapsedTime(&t, ev[s2], ev[s2 + 1]);
      total[s2] += t;
    }
  }
  ctx->have_logits = false;
  for (auto &e : ev) cudaEventDestroy(e);
  ms[0] = float(total[0] / iters);
  ms[1] = float(total[1] / iters);
  if (fused) *fused = took_fused ? 1 : 0;
  return rc;
} FDNN_CATCH

// ---- full forward over host buffers -------------------------------------------------------------

// How a call of n frames is cut over the devices of a group: contiguous shards in device order, whole tiles of 128 frames
// (the row block of every layer kernel), sizes differing by at most one tile, no device used for less than a tile; the
// last shard takes the ragged end.  Host-only (also what tests/test_multirank.py checks without a GPU).
int fdnn_shard_plan(int n_frames, int n_devices, int *first, int *count) {
  if (n_frames < 0 || n_devices <= 0 || !first || !count) return FDNN_EINVAL;
  const int units = (n_frames + 127) / 128;
  const int used = std::max(1, std::min(n_devices, units));
  int f = 0;
  for (int d = 0; d < n_devices; ++d) {
    first[d] = f;
    count[d] = 0;
    if (d < used) {
      const int u = units / used + (d < units % used ? 1 : 0);
      count[d] = std::min(n_frames - f, u * 128);
      f += count[d];
    }
  }
  return used;
}

namespace {

int chunk_frames() {
  static int v = [] {
    const char *e = std::getenv("FDNN_CHUNK_FRAMES");
    int c = e ? std::atoi(e) : 0;
    return c > 0 ? c : 4096;
  }();
  return v;
}

constexpr int kStagedChunk = 512;  // frames per pass when results go through the page-locked staging buffers
constexpr int kSubRows = 128;  // rows per result sub-chunk on the staged (pageable) path: 4 MB of scores at 8000 outputs

// Workspace sizes come in buckets (128 · 2^k frames up to the streaming chunk): callers with variable-length utterances
// then reuse a handful of pooled contexts instead of allocating one per distinct frame count.
int bucket_cap(int n) {
  const int chunk = chunk_frames();
  if (n >= chunk) return chunk;
  int c = 128;
  while (c < n) c *= 2;
  return std::min(c, chunk);
}

// smallest pooled context that holds `cap` frames (none larger than 4× what is needed: a 4096-frame workspace should
// not be tied up by 100-frame calls)
fdnn_ctx *pool_take(fdnn_model *m, int cap) {
  std::lock_guard<std::mutex> lk(m->pool_mu);
  size_t best = m->pool.size();
  for (size_t i = 0; i < m->pool.size(); ++i) {
    const int c = m->pool[i]->cap;
    if (c >= cap && c <= 4 * cap && (best == m->pool.size() || c < m->pool[best]->cap)) best = i;
  }
  if (best == m->pool.size()) return nullptr;
  fdnn_ctx *c = m->pool[best];
  m->pool.erase(m->pool.begin() + long(best));
  return c;
}

void pool_give(fdnn_model *m, fdnn_ctx *c) {
  fdnn_ctx *evict = nullptr;
  {
    std::lock_guard<std::mutex> lk(m->pool_mu);
    m->pool.push_back(c);
    if (m->pool.size() > 16) {
      evict = m->pool.front();  // least recently returned
      m->pool.erase(m->pool.begin());
    }
  }
  destroy_ctx(evict);  // outside pool_mu; takes cuda_mu for the frees only
}

// Copy out of the page-locked transfer buffer into caller memory with non-temporal stores: the destination is megabytes of
// memory nobody will read before the caller does, so ordinary stores would first READ every destination line into the cache
// (read-for-ownership) and evict what the caller's other threads are using.  Source lines were just written by DMA: not
// cached either way.  (Measured on the B200 host: profiles/r2_experiments.md.)
void stream_copy(void *dst, const void *src, size_t bytes) {
  uint8_t *d = static_cast<uint8_t *>(dst);
  const uint8_t *s = static_cast<const uint8_t *>(src);
  const size_t head = (16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15;
  if (bytes < 4096 || head > bytes) {
    std::memcpy(d, s, bytes);
    return;
  }
  std::memcpy(d, s, head);
  d += head;
  s += head;
  bytes -= head;
  const size_t blocks = bytes / 64;
  for (size_t i = 0; i < blocks; ++i) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s) + 0), b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s) + 1),
                  c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s) + 2), e = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s) + 3);
    _mm_stream_si128(reinterpret_cast<__m128i *>(d) + 0, a);
    _mm_stream_si128(reinterpret_cast<__m128i *>(d) + 1, b);
    _mm_stream_si128(reinterpret_cast<__m128i *>(d) + 2, c);
    _mm_stream_si128(reinterpret_cast<__m128i *>(d) + 3, e);
    s += 64;
    d += 64;
  }
  _mm_sfence();
  std::memcpy(d, s, bytes - blocks * 64);
}

bool host_pinned(const void *p) {
  cudaPointerAttributes attr{};
  const bool pinned = cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  return pinned;
}

// How a host thread waits for its results.  Spinning (cudaStreamSynchronize's default) keeps a core busy per caller — with
// several callers per GPU and eight GPUs that starves the callers that have work to do; the driver's blocking sync sleeps on
// an interrupt, which on virtualised hosts wakes up hundreds of microseconds late (measured here: +0.5 ms per call).
// So: poll the event, yielding the core between polls with a short sleep once the wait has lasted longer than a kernel
// launch.  FDNN_SYNC=spin polls without sleeping (lowest latency for one caller on an idle host).
bool sync_by_spinning() {
  static const bool spin = [] {
    const char *e = std::getenv("FDNN_SYNC");
    return e && std::string(e) == "spin";
  }();
  return spin;
}

cudaError_t wait_event(cudaEvent_t ev) {
  const bool spin = sync_by_spinning();
  const auto t0 = std::chrono::steady_clock::now();
  for (;;) {
    const cudaError_t e = cudaEventQuery(ev);
    if (e != cudaErrorNotReady) return e;
    if (!spin && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(20)) {
      timespec ts{0, 25000};  // 25 us; the kernel's timer slack makes it ≈ 80 us
      nanosleep(&ts, nullptr);
    }
  }
}

int ensure_events(fdnn_ctx *x, size_t count) {
  while (x->events.size() < count) {
    cudaEvent_t e;
    CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    x->events.push_back(e);
  }
  return FDNN_OK;
}

// One device's share of a fdnn_calculate call: frames [f0, f0 + n), processed in chunks of at most `cap` on two pooled
// contexts (two streams: while one chunk computes, the next one's input goes up and the previous result comes down).
struct Share {
  fdnn_model *m = nullptr;
  int f0 = 0, n = 0, cap = 0, n_chunks = 0, n_slots = 0;
  fdnn_ctx *slot[2] = {nullptr, nullptr};
  int enqueued = 0, drained = 0;
};

struct CalcCall {
  const float *in;
  bool in_pinned;
  float *out;        // direct destination (may be null when sink is set)
  bool out_direct;   // D2H straight into `out` (page-locked caller memory)
  fdnn_sink_fn sink;
  void *user;
  int I, O;
};

int enqueue_chunk(const CalcCall &call, Share &s, int c) {
  fdnn_ctx *x = s.slot[c % s.n_slots];
  const int f = s.f0 + c * s.cap, m = std::min(s.cap, s.n - c * s.cap);
  const size_t I = size_t(call.I), O = size_t(call.O);
  CUDA_TRY(cudaSetDevice(s.m->device));
  const int subs = call.out_direct ? 0 : (m + kSubRows - 1) / kSubRows;
  if (int rc = ensure_events(x, size_t(1 + subs))) return rc;
  const float *src = call.in + size_t(f) * I;
  if (!call.in_pinned) {
    if (!x->h_in) CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&x->h_in), size_t(x->cap) * I * 4, cudaHostAllocPortable));
    std::memcpy(x->h_in, src, size_t(m) * I * 4);
    src = x->h_in;
  }
  CUDA_TRY(cudaMemcpyAsync(x->d_in, src, size_t(m) * I * 4, cudaMemcpyHostToDevice, x->stream));
  if (int rc = run_pass(x, x->d_in, m, x->d_logits, true, x->stream)) return rc;
  x->have_logits = false;
  if (call.out_direct) {
    CUDA_TRY(cudaMemcpyAsync(call.out + size_t(f) * O, x->d_logits, size_t(m) * O * 4, cudaMemcpyDeviceToHost, x->stream));
  } else {
    if (!x->h_out) CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&x->h_out), size_t(x->cap) * O * 4, cudaHostAllocPortable));
    for (int k = 0; k < subs; ++k) {
      const int r0 = k * kSubRows, rows = std::min(kSubRows, m - r0);
      CUDA_TRY(cudaMemcpyAsync(x->h_out + size_t(r0) * O, x->d_logits + size_t(r0) * O, size_t(rows) * O * 4, cudaMemcpyDeviceToHost, x->stream));
      CUDA_TRY(cudaEventRecord(x->events[size_t(1 + k)], x->stream));
    }
  }
  CUDA_TRY(cudaEventRecord(x->events[0], x->stream));
  return FDNN_OK;
}

int drain_chunk(const CalcCall &call, Share &s, int c) {
  fdnn_ctx *x = s.slot[c % s.n_slots];
  const int f = s.f0 + c * s.cap, m = std::min(s.cap, s.n - c * s.cap);
  const size_t O = size_t(call.O);
  CUDA_TRY(cudaSetDevice(s.m->device));
  if (!call.out_direct) {
    const int subs = (m + kSubRows - 1) / kSubRows;
    for (int k = 0; k < subs; ++k) {
      const int r0 = k * kSubRows, rows = std::min(kSubRows, m - r0);
      CUDA_TRY(wait_event(x->events[size_t(1 + k)]));
      if (call.sink) {
        if (call.sink(call.user, f + r0, rows, x->h_out + size_t(r0) * O) != 0) {
          set_error("the result sink reported a failure");
          return FDNN_EINVAL;
        }
      } else {
        stream_copy(call.out + size_t(f + r0) * O, x->h_out + size_t(r0) * O, size_t(rows) * O * 4);
      }
    }
  }
  CUDA_TRY(wait_event(x->events[0]));
  return FDNN_OK;
}

int calculate_impl(fdnn_model *model, const float *in, int n, int dim, float *out, fdnn_sink_fn sink, void *user) {
  if (!model || n < 0) {
    set_error("bad argument to fdnn_calculate");
    return FDNN_EINVAL;
  }
  if (dim != model->hdr.in_dim) {  // QuantizedDnn.java:157-161
    set_error("input dimension " + std::to_string(dim) + " does not match the network's " + std::to_string(model->hdr.in_dim));
    return FDNN_EINVAL;
  }
  if (n == 0) return FDNN_OK;  // QuantizedDnn.java:154-156
  if (!in || (!out && !sink)) {
    set_error("null buffer");
    return FDNN_EINVAL;
  }
  DeviceGuard g(model->device);
  if (!g.ok) {
    set_error("cudaSetDevice failed");
    return FDNN_ECUDA;
  }
  CalcCall call{};
  call.in = in;
  // FDNN_STAGE=0 (experiments): pageable caller memory straight into cudaMemcpyAsync (the driver stages it, synchronously)
  static const bool stage = env_flag("FDNN_STAGE", true);
  call.in_pinned = !stage || host_pinned(in);
  call.out = out;
  call.out_direct = sink == nullptr && (!stage || host_pinned(out));
  call.sink = sink;
  call.user = user;
  call.I = model->hdr.in_dim;
  call.O = model->hdr.out_dim;

  // one contiguous shard per device of the group (frames are independent: no collective), whole tiles of 128 frames each
  const int n_dev = model->group.empty() ? 1 : int(model->group.size());
  std::vector<int> first(size_t(n_dev), 0), count(size_t(n_dev), 0);
  const int used = fdnn_shard_plan(n, n_dev, first.data(), count.data());
  std::vector<Share> shares{size_t(used)};
  int rc = FDNN_OK;
  for (int d = 0; d < used; ++d) {
    Share &s = shares[size_t(d)];
    s.m = model->group.empty() ? model : model->group[size_t(d)];
    s.f0 = first[size_t(d)];
    s.n = count[size_t(d)];
    // Staged (pageable) callers: the calling thread's copy-out, not the GPU, is the slow stage (≈ 10-15 GB/s per host thread), and
    // it runs at its best out of transfer buffers small enough to still sit in the last-level cache the DMA wrote them to — chunks
    // of 512 frames (16 MB of scores), which also lets upload, compute, download and copy-out of one call overlap.
    s.cap = call.out_direct ? bucket_cap(s.n) : std::min(bucket_cap(s.n), kStagedChunk);
    s.n_chunks = (s.n + s.cap - 1) / s.cap;
    s.n_slots = s.n_chunks > 1 ? 2 : 1;
    for (int k = 0; k < s.n_slots && rc == FDNN_OK; ++k) {
      s.slot[k] = pool_take(s.m, s.cap);
      if (!s.slot[k]) rc = create_ctx(s.m, s.cap, &s.slot[k]);
    }
  }
  // One host thread drives every device: all work is asynchronous, so the calling thread enqueues round-robin and then
  // collects in the same order; the sink (JNI: SetFloatArrayRegion) is only ever called from the calling thread.
  int max_chunks = 0;
  for (const Share &s : shares) max_chunks = std::max(max_chunks, s.n_chunks);
  for (int c = 0; c < max_chunks + 2 && rc == FDNN_OK; ++c) {
    for (Share &s : shares) {
      // a slot is reused by chunk c only after chunk c − 2 has been collected (its staging buffers are free again)
      if (c >= 2 && c - 2 < s.n_chunks && rc == FDNN_OK) {
        rc = drain_chunk(call, s, c - 2);
        s.drained = c - 1;
      }
      if (c < s.n_chunks && rc == FDNN_OK) {
        rc = enqueue_chunk(call, s, c);
        s.enqueued = c + 1;
      }
    }
  }
  for (Share &s : shares) {
    if (rc != FDNN_OK)  // let whatever was enqueued finish before the workspaces go away
      for (int k = 0; k < s.n_slots; ++k)
        if (s.slot[k]) {
          cudaSetDevice(s.m->device);
          cudaStreamSynchronize(s.slot[k]->stream);
        }
    for (int k = 0; k < s.n_slots; ++k)
      if (s.slot[k]) {
        if (rc == FDNN_OK)
          pool_give(s.m, s.slot[k]);
        else
          destroy_ctx(s.slot[k]);
      }
  }
  if (rc != FDNN_OK) cudaGetLastError();
  return rc;
}

}  // namespace

// All frames of the context at once (BASELINE config 3): masks up, one masked-softmax launch, scores down.  With page-locked caller
// memory all three are asynchronous and the thread naps until they are done; pageable memory goes through the driver's own
// staging (measured on B200: faster than staging it here — 152 k against 123 k frames/s at batch 512, profiles/r2_experiments.md).
int fdnn_ctx_lazy_batch(fdnn_ctx *ctx, const int8_t *masks, float *out) try {
  if (!ctx || !masks || !out) {
    set_error("null argument");
    return FDNN_EINVAL;
  }
  if (!ctx->have_logits) {
    set_error("calculateUntilOutput has not been run on this context");
    return FDNN_EINVAL;
  }
  DeviceGuard g(ctx->model->device);
  const size_t O = size_t(ctx->model->hdr.out_dim);
  const int n = ctx->last_frames;
  if (!ctx->d_masks) CUDA_TRY(cudaMalloc(&ctx->d_masks, size_t(ctx->cap) * O));
  // the masked softmax must not overwrite the resident logits (later lazy calls need them)
  if (!ctx->d_lazy) CUDA_TRY(cudaMalloc(&ctx->d_lazy, size_t(ctx->cap) * O * 4));
  CUDA_TRY(cudaMemcpyAsync(ctx->d_masks, masks, size_t(n) * O, cudaMemcpyHostToDevice, ctx->stream));
  if (int rc = enqueue_softmax(ctx, ctx->d_logits, ctx->d_masks, n, ctx->d_lazy, ctx->stream)) return rc;
  CUDA_TRY(cudaMemcpyAsync(out, ctx->d_lazy, size_t(n) * O * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (int rc = ensure_events(ctx, 1)) return rc;
  CUDA_TRY(cudaEventRecord(ctx->events[0], ctx->stream));
  CUDA_TRY(wait_event(ctx->events[0]));
  return FDNN_OK;
} FDNN_CATCH

int fdnn_calculate(fdnn_model *model, const float *in, int n, int dim, int batch_hint, float *out) try {
  (void) batch_hint;
  return calculate_impl(model, in, n, dim, out, nullptr, nullptr);
} FDNN_CATCH

int fdnn_calculate_sink(fdnn_model *model, const float *in, int n, int dim, fdnn_sink_fn sink, void *user) try {
  if (!sink) {
    set_error("null sink");
    return FDNN_EINVAL;
  }
  return calculate_impl(model, in, n, dim, nullptr, sink, user);
} FDNN_CATCH

// ---- misc ---------------------------------------------------------------------------------------------

int fdnn_host_alloc(void **ptr, size_t bytes) {
  if (!ptr) return FDNN_EINVAL;
  int dev = 0;
  if (int rc = usable_device(-1, &dev)) return rc;
  CUDA_TRY(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable));
  return FDNN_OK;
}

int fdnn_host_free(void *ptr) {
  if (ptr) CUDA_TRY(cudaFreeHost(ptr));
  return FDNN_OK;
}

long long fdnn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
