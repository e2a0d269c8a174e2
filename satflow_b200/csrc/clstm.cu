// libclstm.so — host side of the C ABI declared in include/clstm.h.
//
// Plans own nothing but descriptors: the caller binds one device workspace, the plan carves it
// into the HBM-resident tensors of the rollout (DESIGN.md "Data layout in HBM"), encodes the TMA
// tensor maps once, and every entry point only enqueues kernels on the caller's stream.
//
// Reference lines each sequence follows are cited next to it (paths relative to the reference
// repository root, openclimatefix/satflow v0.3.36).
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <type_traits>
#include <vector>

#include "../../include/clstm.h"
#include "convgemm.cuh"
#include "dgradT.cuh"
#include "dgradT_fused2.cuh"
#include "head_rows.cuh"
#include "cellstep_pair.cuh"
#include "pointwise.cuh"
#include "rollout_persist.cuh"
#include "wgrad.cuh"

using namespace clstm;

namespace {

thread_local char g_err[1024] = "";
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CU_TRY(expr)                                                                                  \
  do {                                                                                                \
    cudaError_t e__ = (expr);                                                                         \
    if (e__ != cudaSuccess)                                                                           \
      return fail(CLSTM_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
  } while (0)

#define RC_TRY(expr)            \
  do {                          \
    int rc__ = (expr);          \
    if (rc__ != 0) return rc__; \
  } while (0)

// In-situ launch trace (clstm_trace_enable / clstm_trace_report): one CUDA event after every launch on the stream of
// the current API call, so per-kernel durations can be read under the power state of the real step — isolated
// kernel timings and the step disagree on a power-capped B200 (DESIGN.md finding 9).  Off by default.
struct LaunchTrace {
  bool on = false;
  cudaStream_t st = nullptr;
  std::vector<cudaEvent_t> ev;
  std::vector<const char*> name;  // nullptr = start of an API call (the interval before it is time outside the library)
  size_t n = 0;
  void mark(const char* what) {
    if (!on || n >= ev.size()) return;
    if (cudaEventRecord(ev[n], st) != cudaSuccess) return;
    name[n++] = what;
  }
};
LaunchTrace g_trace;

inline void trace_begin(void* stream) {
  if (!g_trace.on) return;
  g_trace.st = static_cast<cudaStream_t>(stream);
  g_trace.mark(nullptr);
}

inline int after_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) return fail(CLSTM_ECUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  if (g_trace.on) g_trace.mark(what);
  return 0;
}

inline int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

// Schedule knobs.  Every setting computes the same function (tests/test_gpu_parity.py::test_alternative_schedules_*
// keeps each one parity-green); they exist for same-session A/B measurements (tools/ab_fuse.py).  They are read from
// the environment ONCE, when a plan is created, never at launch time; nothing here can change a result.
struct Knobs {
  int stages = kMaxStages;  // CLSTM_STAGES: cap on the operand ring depth of the conv GEMM
  int rotate = 1;           // CLSTM_ROTATE: per-tile rotated K loop in the conv GEMM
  int rotate_d = 0;         // CLSTM_ROTATE_D: the same in dgradT (measured: no gain)
  int staged = 1;           // CLSTM_STAGED: shared-memory staged TMA-store epilogue (0: direct per-thread stores;
                            // 2: 8-channel groups + c_prev in registers -> a fourth operand stage fits)
  int dgradT = 1;           // CLSTM_DGRADT: transposed dgrad (0: pixel-major dgrad through convgemm)
  int fuse_gate = 1;        // CLSTM_FUSE_GATE: gate gradient fused into the dgrad epilogue
  int fuse_pf = 1;          // CLSTM_FUSE_PF: its L2 prefetch distance in groups
  int fuse_sets = 4;        // CLSTM_FUSE_SETS: register sets of loads in flight per epilogue thread (2; 4 = setmaxnreg)
  int fuse_workers = 2;     // CLSTM_FUSE_WORKERS: gate gradient on dedicated worker warps (dgradT_fused2.cuh) with this
                            // 2 register sets of loads in flight (2); 0 = the epilogue warps do it themselves.  (A 4-set
                            // variant spilled at 136 registers and measured 1196 us: removed.)
  int head_rows = 1;        // CLSTM_HEAD_ROWS: row-marching output head (0: implicit-GEMM head)
  int head_band = 32;       // CLSTM_HEAD_BAND: rows per band of the row-marching head
  int wg_halo = 1;          // CLSTM_WG_HALO: halo-row wgrad
  int wg_gate = 0;          // CLSTM_WG_GATE: gate gradient on worker warps inside wgrad (1: plain, 2: setmaxnreg)
  int hybrid_pct = 0;       // CLSTM_HYBRID: percent of the pixels whose gate gradient stays in the dgrad epilogue, the rest
                            // runs on worker warps of the following wgrad launch (0 = off: all in the dgrad epilogue)
  int hybrid_wg = 2;        // CLSTM_HYBRID_WG: which wgrad worker variant the hybrid schedule uses (1 or 2)
  int wg_group = kWgMaxGroupBlocks;  // CLSTM_WG_GROUP: column blocks per wgrad CTA
  int overlap = 0;          // CLSTM_OVERLAP: wgrad on a side stream
  int recomp_c = 1;         // CLSTM_RECOMP_C: the fused gate gradient rebuilds c' = f c + i g from the saved gates instead of
                            // reading it (fp16 operands only: -8 % of the fused dgrad launch's HBM bytes)
  int head_fuse = 1;        // CLSTM_HEAD_FUSE: the head's dgrad of frame t rides as a second K segment inside the fused dgrad
                            // launch whose epilogue runs the top decoder cell's gate gradient of step t (no dstack)
  int state16 = 1;          // CLSTM_STATE16: dc and the cells' own dh_prev in 16 bits (at the dz scale) between the launches of the fused
                            // backward chain (fp16 operands, worker-warp fused dgrad with recomputed c' only)
  int c16 = 1;              // CLSTM_C16: the cell state c crosses HBM in 16 bits (x 2^8) in big (non-persistent) fp16 rollouts
  int pair = 1;             // CLSTM_PAIR: cell step on CTA pairs (cta_group::2, cellstep_pair.cuh) for shapes with at least two
                            // waves of tiles
  int persist = 1;          // CLSTM_PERSIST: one persistent launch for the whole forward chain when the state fits on chip
  int graph = 1;            // CLSTM_GRAPH: launch-bound (small) rollouts replay their forward / backward as CUDA graphs
  void read() {
    stages = env_int("CLSTM_STAGES", stages);
    rotate = env_int("CLSTM_ROTATE", rotate);
    rotate_d = env_int("CLSTM_ROTATE_D", rotate_d);
    staged = env_int("CLSTM_STAGED", staged);
    dgradT = env_int("CLSTM_DGRADT", dgradT);
    fuse_gate = env_int("CLSTM_FUSE_GATE", fuse_gate);
    fuse_pf = env_int("CLSTM_FUSE_PF", fuse_pf);
    fuse_sets = env_int("CLSTM_FUSE_SETS", fuse_sets);
    fuse_workers = env_int("CLSTM_FUSE_WORKERS", fuse_workers);
    head_rows = env_int("CLSTM_HEAD_ROWS", head_rows);
    head_band = env_int("CLSTM_HEAD_BAND", head_band);
    wg_halo = env_int("CLSTM_WG_HALO", wg_halo);
    wg_gate = env_int("CLSTM_WG_GATE", wg_gate);
    hybrid_pct = env_int("CLSTM_HYBRID", hybrid_pct);
    hybrid_wg = env_int("CLSTM_HYBRID_WG", hybrid_wg);
    wg_group = env_int("CLSTM_WG_GROUP", wg_group);
    if (wg_group < 1 || wg_group > kWgMaxGroupBlocks) wg_group = kWgMaxGroupBlocks;
    overlap = env_int("CLSTM_OVERLAP", overlap);
    recomp_c = env_int("CLSTM_RECOMP_C", recomp_c);
    head_fuse = env_int("CLSTM_HEAD_FUSE", head_fuse);
    state16 = env_int("CLSTM_STATE16", state16);
    c16 = env_int("CLSTM_C16", c16);
    pair = env_int("CLSTM_PAIR", pair);
    persist = env_int("CLSTM_PERSIST", persist);
    graph = env_int("CLSTM_GRAPH", graph);
  }
};

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// --------------------------------------------------------------------------- CUDA-graph replay of launch-bound calls
// A small rollout (BASELINE configs[0]: 64x64, batch 2) is ~75 launches of a few microseconds each per backward: the
// host cannot enqueue them as fast as the GPU retires them (measured: 1.5 ms of enqueue for 1.8 ms of step).  The
// sequence of launches of one API call depends only on the plan and on the pointers passed in, so it is captured once
// per distinct pointer set (stream capture of the very same host code) and replayed with one cudaGraphLaunch.
// PyTorch's caching allocator hands a training loop the same blocks every iteration, so the pointer sets repeat.
// A key is captured the SECOND time it is seen (one-off calls never pay for an instantiation); at most kMaxGraphs
// executables are kept per call site (LRU); any capture failure disables the cache for the plan and runs eagerly.
struct GraphCache {
  static constexpr size_t kMaxGraphs = 4;
  struct Entry {
    std::vector<const void*> key;
    cudaGraphExec_t exec = nullptr;
    uint64_t last = 0;
    uint64_t launches = 0;  // kernels in the graph (keeps clstm_launch_count meaningful across replays)
  };
  std::vector<Entry> entries;
  std::vector<std::vector<const void*>> seen;  // keys run eagerly once (small ring)
  uint64_t tick = 0, replays = 0, captures = 0;
  bool disabled = false;
  // Capture happens on a private stream: the caller's stream may be the legacy default stream (PyTorch's current
  // stream usually is), which cannot be captured; the instantiated graph is then launched on the caller's stream.
  cudaStream_t capture_stream = nullptr;
  void clear() {
    for (Entry& e : entries)
      if (e.exec) cudaGraphExecDestroy(e.exec);
    entries.clear();
    seen.clear();
    if (capture_stream) cudaStreamDestroy(capture_stream);
    capture_stream = nullptr;
  }
};

template <typename F>
int run_graphed(GraphCache& gc, bool enabled, std::vector<const void*> key, cudaStream_t st, F&& body) {
  if (!enabled || gc.disabled || g_trace.on) return body(st);
  ++gc.tick;
  for (GraphCache::Entry& e : gc.entries)
    if (e.key == key) {
      e.last = gc.tick;
      ++gc.replays;
      CU_TRY(cudaGraphLaunch(e.exec, st));
      g_launches.fetch_add(e.launches, std::memory_order_relaxed);
      return 0;
    }
  bool seen_before = false;
  for (const auto& k : gc.seen) seen_before = seen_before || (k == key);
  if (!seen_before) {
    if (gc.seen.size() >= 8) gc.seen.erase(gc.seen.begin());
    gc.seen.push_back(key);
    return body(st);
  }
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return body(st);  // the caller is capturing this stream itself
  }
  if (!gc.capture_stream && cudaStreamCreateWithFlags(&gc.capture_stream, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    gc.disabled = true;
    return body(st);
  }
  if (cudaStreamBeginCapture(gc.capture_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    gc.disabled = true;
    return body(st);
  }
  const uint64_t l0 = g_launches.load();
  const int rc = body(gc.capture_stream);
  const uint64_t captured_launches = g_launches.load() - l0;
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(gc.capture_stream, &graph);
  cudaGraphExec_t exec = nullptr;
  if (rc != 0 || ce != cudaSuccess || graph == nullptr || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    gc.disabled = true;
    return rc != 0 ? rc : body(st);  // nothing was executed during the capture
  }
  cudaGraphDestroy(graph);
  if (gc.entries.size() >= GraphCache::kMaxGraphs) {
    size_t lru = 0;
    for (size_t i = 1; i < gc.entries.size(); ++i)
      if (gc.entries[i].last < gc.entries[lru].last) lru = i;
    cudaGraphExecDestroy(gc.entries[lru].exec);
    gc.entries.erase(gc.entries.begin() + static_cast<long>(lru));
  }
  GraphCache::Entry e;
  e.key = std::move(key), e.exec = exec, e.last = gc.tick, e.launches = captured_launches;
  gc.entries.push_back(std::move(e));
  ++gc.captures;
  CU_TRY(cudaGraphLaunch(exec, st));
  return 0;
}

// --------------------------------------------------------------------------- device / driver
struct DeviceInfo {
  int ordinal = -1;
  int sms = 0;
  int smem_optin = 0;
};

int get_device(DeviceInfo* d) {
  int dev = 0;
  CU_TRY(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return fail(CLSTM_ENODEV, "device %d (%s) has compute capability %d.%d; this library is sm_100a only", dev,
                prop.name, prop.major, prop.minor);
  d->ordinal = dev;
  d->sms = prop.multiProcessorCount;
  d->smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode(EncodeTiledFn* out) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CU_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !p) return fail(CLSTM_ECUDA, "cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  *out = fn;
  return 0;
}

// NHWC activation tensor [N][H][W][C] of 16-bit elements; box = 64 channels x boxW x boxH x 1 image,
// 128-byte swizzle, out-of-bounds (including negative coordinates) reads as zero == conv padding.
int make_map_act(CUtensorMap* m, int dtype, const void* ptr, int C, int W, int H, long long N, int boxW, int boxH) {
  EncodeTiledFn enc;
  RC_TRY(get_encode(&enc));
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)boxW, (cuuint32_t)boxH, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, dtype == CLSTM_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                   const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(CLSTM_ECUDA, "cuTensorMapEncodeTiled(act C=%d W=%d H=%d N=%lld box %dx%d) -> %d", C, W, H, N, boxW,
                boxH, (int)r);
  return 0;
}

// Epilogue tensor maps: [N][H][W][C] tensors accessed in [boxH x boxW pixel] x 16-channel boxes.  fp32 tensors
// (c states, dgrad outputs) use 64-byte rows + SWIZZLE_64B, 16-bit tensors (h, gates) 32-byte rows + SWIZZLE_32B,
// matching the conflict-free staging layout of the convgemm epilogue.  TMA stores clip out-of-bounds pixels.
int make_map_epi(CUtensorMap* m, int elem_bytes, int dtype16, const void* ptr, int C, int W, int H, long long N,
                 int boxW, int boxH, int box_c = 16) {
  EncodeTiledFn enc;
  RC_TRY(get_encode(&enc));
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * elem_bytes, (cuuint64_t)W * C * elem_bytes,
                           (cuuint64_t)H * W * C * elem_bytes};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)boxW, (cuuint32_t)boxH, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  const int row_bytes = box_c * elem_bytes;
  const CUtensorMapSwizzle sw =
      row_bytes >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                       : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                          : (row_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE));
  CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                           : (dtype16 == CLSTM_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                                                   : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  CUresult r = enc(m, dt, 4, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(CLSTM_ECUDA, "cuTensorMapEncodeTiled(epilogue C=%d W=%d H=%d N=%lld elem %d) -> %d", C, W, H, N,
                elem_bytes, (int)r);
  return 0;
}

// fp32 [N][H][W][C] tensor written in [16 px along W] x 64-channel boxes (256-byte rows, no swizzle): the
// transposed dgrad's output blocks.
int make_map_out64(CUtensorMap* m, const void* ptr, int C, int W, int H, long long N, int dtype16 = -1) {
  // dtype16 < 0: fp32 elements; CLSTM_F16 / CLSTM_BF16: 16-bit elements (the 16-bit dh_prev of the default backward schedule)
  EncodeTiledFn enc;
  RC_TRY(get_encode(&enc));
  const cuuint64_t eb = dtype16 < 0 ? 4 : 2;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * eb, (cuuint64_t)W * C * eb, (cuuint64_t)H * W * C * eb};
  cuuint32_t box[4] = {64, 16, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  const CUtensorMapDataType dt = dtype16 < 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : (dtype16 == CLSTM_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  CUresult r = enc(m, dt, 4, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(CLSTM_ECUDA, "cuTensorMapEncodeTiled(out64 C=%d W=%d H=%d) -> %d", C, W, H, (int)r);
  return 0;
}

// Packed weight matrix [rows][K] (K-major); box = 64 k x boxRows rows.
int make_map_w(CUtensorMap* m, int dtype, const void* ptr, int K, int rows, int boxRows) {
  EncodeTiledFn enc;
  RC_TRY(get_encode(&enc));
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)boxRows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, dtype == CLSTM_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(CLSTM_ECUDA, "cuTensorMapEncodeTiled(weights K=%d rows=%d box %d) -> %d", K, rows, boxRows, (int)r);
  return 0;
}

// --------------------------------------------------------------------------- geometry
struct Geo {
  int B, H, W;
  int BW, BH, tiles_w, tiles_h;      // 128-pixel tiles (pixel-major GEMMs)
  int BW2, BH2, tiles_w2, tiles_h2;  // 64-pixel tiles (wgrad K steps)
  size_t npix() const { return static_cast<size_t>(B) * H * W; }
};

Geo make_geo(int B, int H, int W) {
  Geo g;
  g.B = B, g.H = H, g.W = W;
  int bw = 8;
  while (bw < W && bw < 128) bw <<= 1;
  g.BW = bw, g.BH = 128 / bw;
  g.tiles_w = (W + g.BW - 1) / g.BW, g.tiles_h = (H + g.BH - 1) / g.BH;
  g.BW2 = bw > 64 ? 64 : bw, g.BH2 = 64 / g.BW2;
  g.tiles_w2 = (W + g.BW2 - 1) / g.BW2, g.tiles_h2 = (H + g.BH2 - 1) / g.BH2;
  return g;
}

int pad_hidden(int hid) {
  if (hid <= 64) return 64;
  if (hid <= 128) return 128;
  if (hid <= 256) return 256;
  if (hid <= 512) return 512;
  return -1;
}

const int kPackBlocks = 148 * 8;
const int kGateGradBlocks = 148 * 4;
const int kBiasRowsMax = 148 * 16;  // bias partial rows: one per gate-worker warp of a wgrad CTA (>= kGateGradBlocks)

// Everything a cell step needs besides the cell itself.
struct Ctx {
  DeviceInfo dev;
  Geo geo;
  int dtype = CLSTM_F16;
  int HP = 0;
  int training = 0;
  float grad_scale = 0.f;
  bool c16 = false;        // rollout plans: the c stacks hold E values times kCScale (CLSTM_C16), see decide_c16
  void* dz = nullptr;      // E [npix][4HP]  (buffer 0; == dzb[0])
  void* dzb[2] = {nullptr, nullptr};  // two dz buffers: gate-grad of step n+1 overlaps the wgrad of step n
  float* scale = nullptr;  // device {S, 1/S}
  unsigned int* amax = nullptr;
  Knobs knobs;
  CUtensorMap m_dz128, m_dz64;
  CUtensorMap m_dz128b[2], m_dz64b[2];
  cudaStream_t side = nullptr;                    // weight-gradient stream (created at bind)
  cudaEvent_t ev_d[2] = {nullptr, nullptr};       // dgrad of the step using buffer i has been enqueued/finished
  cudaEvent_t ev_w[2] = {nullptr, nullptr};       // wgrad has finished reading buffer i
  cudaEvent_t ev_fork = nullptr;
};

struct CellState {
  CellGeom g;
  int T = 0;       // steps this cell runs per rollout
  int Kf = 0;      // forward K
  int Kd = 0;      // dgrad K
  int rows_d = 0;  // dgrad output columns
  int with_x = 0;
  int n_tile_d = 0;
  int slots_h = 0, slots_c = 0;
  int wg_total = 0, wg_group = 0, wg_splits = 0;
  bool bwd_started = false;
  // packed parameters
  void* wp = nullptr;
  float* bias_p = nullptr;
  void* wd = nullptr;
  // state stacks
  void* h = nullptr;      // E [slots_h][npix][HP]
  float* c = nullptr;     // fp32 [slots_c][npix][HP]
  void* gates = nullptr;  // E [T][npix][4HP] (training)
  // backward scratch
  float* dh_own = nullptr;  // fp32 [npix][HP]
  float* dxb = nullptr;     // fp32 [npix][CIP]
  float* dc = nullptr;      // fp32 [npix][HP]
  float* wpart = nullptr;   // fp32 [splits][4HP][Kf]
  float* bpart = nullptr;   // fp32 [kGateGradBlocks][4HP]
  CUtensorMap m_h128, m_h64, m_wp, m_wp128, m_wd;  // m_wp128: 128-row weight boxes (a CTA pair's halves)
  CUtensorMap m_h66;                                // wgrad halo rows: box 64 ch x 66 px x 1 row
  CUtensorMap m_wdT, m_dxT, m_dhT, m_dhT16;         // transposed dgrad: 128-row weight boxes, 64-channel output boxes
                                                    // (m_dhT16: dh_own viewed as 16-bit elements, CLSTM_STATE16)
  CUtensorMap m_c16, m_h16, m_g16, m_dh16, m_dx16;  // epilogue (staged store / c_prev load) maps
  CUtensorMap m_c8, m_h8, m_g8;                     // the same with 8-channel boxes (CLSTM_STAGED=2)

  size_t h_slot_elems(const Geo& geo) const { return geo.npix() * g.HP; }
};

// Where a cell's input comes from at one step.
struct InputRef {
  const CUtensorMap* map128 = nullptr;
  const CUtensorMap* map64 = nullptr;
  const CUtensorMap* map66 = nullptr;    // wgrad halo rows (conv inputs only)
  int b_off = 0;  // image offset (slot * B)
};

struct Carver {
  uint8_t* base = nullptr;
  size_t off = 0;
  template <typename T>
  T* take(size_t bytes) {
    off = align_up(off, 1024);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += bytes;
    return p;
  }
};

void wgrad_shape(const DeviceInfo& dev, int max_group, int total_blocks, int n_blocks, long long p_tiles,
                 int* group_size, int* splits) {
  int groups = (total_blocks + max_group - 1) / max_group;
  int gs = (total_blocks + groups - 1) / groups;
  groups = (total_blocks + gs - 1) / gs;
  int sms = dev.sms > 0 ? dev.sms : 148;
  long long sp = sms / (groups * n_blocks);
  if (sp > p_tiles) sp = p_tiles;
  if (sp < 1) sp = 1;
  *group_size = gs;
  *splits = static_cast<int>(sp);
}

// Fills the derived fields of a cell.  in_col: the input is an im2col'd tensor (rollout encoder_1).
void init_cell(CellState* cs, const Ctx& ctx, int cin, int hid, int kh, int kw, int in_col, int with_x, int T,
               int in_scaled) {
  CellGeom& g = cs->g;
  g.cin = cin, g.hid = hid, g.HP = ctx.HP, g.kh = kh, g.kw = kw, g.in_col = in_col;
  g.CIP = round_up(cin, 64);
  g.in_scaled = in_scaled;
  g.KIN = in_col ? round_up(kh * kw * cin, 64) : kh * kw * g.CIP;
  cs->T = T;
  cs->Kf = g.KIN + kh * kw * g.HP;
  cs->Kd = kh * kw * 4 * g.HP;
  cs->with_x = with_x;
  cs->rows_d = (with_x ? g.CIP : 0) + g.HP;
  int nt = 256;
  while (cs->rows_d % nt) nt -= 64;
  cs->n_tile_d = nt;
  cs->wg_total = cs->Kf / 64;
  const long long p_tiles = static_cast<long long>(ctx.geo.B) * ctx.geo.tiles_w2 * ctx.geo.tiles_h2;
  wgrad_shape(ctx.dev, ctx.knobs.wg_group, cs->wg_total, 4 * g.HP / 128, p_tiles, &cs->wg_group, &cs->wg_splits);
}

// `sv` receives what a backward needs from the forward (packed weights, state stacks, gates), `sc` the scratch that
// any later call may overwrite.  A rollout plan passes the same carver twice (one workspace); a cell plan keeps the
// two apart so that every autograd node can own its saved region (clstm_cell_plan_bind_split).
void carve_cell(Carver& sv, Carver& sc, CellState& cs, const Ctx& ctx) {
  const size_t npix = ctx.geo.npix();
  const int HP = ctx.HP;
  cs.wp = sv.take<void>(static_cast<size_t>(4 * HP) * cs.Kf * 2);
  cs.bias_p = sv.take<float>(static_cast<size_t>(4 * HP) * 4);
  cs.h = sv.take<void>(static_cast<size_t>(cs.slots_h) * npix * HP * 2);
  cs.c = sv.take<float>(static_cast<size_t>(cs.slots_c) * npix * HP * (ctx.c16 ? 2 : 4));  // 16-bit stacks: half the bytes
  if (ctx.training) {
    cs.wd = sv.take<void>(static_cast<size_t>(cs.rows_d) * cs.Kd * 2);
    cs.gates = sv.take<void>(static_cast<size_t>(cs.T) * npix * 4 * HP * 2);
    cs.dh_own = sc.take<float>(npix * HP * 4);
    cs.dxb = cs.with_x ? sc.take<float>(npix * cs.g.CIP * 4) : nullptr;
    cs.dc = sc.take<float>(npix * HP * 4);
    cs.wpart = sc.take<float>(static_cast<size_t>(cs.wg_splits) * 4 * HP * cs.Kf * 4);
    cs.bpart = sc.take<float>(static_cast<size_t>(kBiasRowsMax) * 4 * HP * 4);
  }
}

int map_cell(CellState& cs, const Ctx& ctx) {
  const Geo& g = ctx.geo;
  const long long imgs = static_cast<long long>(cs.slots_h) * g.B;
  RC_TRY(make_map_act(&cs.m_h128, ctx.dtype, cs.h, ctx.HP, g.W, g.H, imgs, g.BW, g.BH));
  RC_TRY(make_map_act(&cs.m_h64, ctx.dtype, cs.h, ctx.HP, g.W, g.H, imgs, g.BW2, g.BH2));
  if (g.BW2 == 64 && g.BH2 == 1) RC_TRY(make_map_act(&cs.m_h66, ctx.dtype, cs.h, ctx.HP, g.W, g.H, imgs, 66, 1));
  RC_TRY(make_map_w(&cs.m_wp, ctx.dtype, cs.wp, cs.Kf, 4 * ctx.HP, 256));
  RC_TRY(make_map_w(&cs.m_wp128, ctx.dtype, cs.wp, cs.Kf, 4 * ctx.HP, 128));
  if (ctx.training)
    RC_TRY(make_map_w(&cs.m_wd, ctx.dtype, cs.wd, cs.Kd, cs.rows_d, cs.n_tile_d));
  RC_TRY(make_map_epi(&cs.m_c16, ctx.c16 ? 2 : 4, ctx.dtype, cs.c, ctx.HP, g.W, g.H,
                      static_cast<long long>(cs.slots_c) * g.B, g.BW, g.BH));
  RC_TRY(make_map_epi(&cs.m_h16, 2, ctx.dtype, cs.h, ctx.HP, g.W, g.H, imgs, g.BW, g.BH));
  RC_TRY(make_map_epi(&cs.m_c8, 4, ctx.dtype, cs.c, ctx.HP, g.W, g.H, static_cast<long long>(cs.slots_c) * g.B, g.BW, g.BH,
                      8));
  RC_TRY(make_map_epi(&cs.m_h8, 2, ctx.dtype, cs.h, ctx.HP, g.W, g.H, imgs, g.BW, g.BH, 8));
  if (ctx.training)
    RC_TRY(make_map_epi(&cs.m_g8, 2, ctx.dtype, cs.gates, 4 * ctx.HP, g.W, g.H, static_cast<long long>(cs.T) * g.B, g.BW,
                        g.BH, 8));
  if (ctx.training) {
    RC_TRY(make_map_epi(&cs.m_g16, 2, ctx.dtype, cs.gates, 4 * ctx.HP, g.W, g.H, static_cast<long long>(cs.T) * g.B,
                        g.BW, g.BH));
    RC_TRY(make_map_epi(&cs.m_dh16, 4, ctx.dtype, cs.dh_own, ctx.HP, g.W, g.H, g.B, g.BW, g.BH));
    RC_TRY(make_map_w(&cs.m_wdT, ctx.dtype, cs.wd, cs.Kd, cs.rows_d, 128));
    RC_TRY(make_map_out64(&cs.m_dhT, cs.dh_own, ctx.HP, g.W, g.H, g.B));
    RC_TRY(make_map_out64(&cs.m_dhT16, cs.dh_own, ctx.HP, g.W, g.H, g.B, ctx.dtype));
    if (cs.with_x) RC_TRY(make_map_out64(&cs.m_dxT, cs.dxb, cs.g.CIP, g.W, g.H, g.B));
    if (cs.with_x) RC_TRY(make_map_epi(&cs.m_dx16, 4, ctx.dtype, cs.dxb, cs.g.CIP, g.W, g.H, g.B, g.BW, g.BH));
  }
  return 0;
}

// --------------------------------------------------------------------------- launch helpers
template <typename E, int EPI>
int launch_convgemm(const Ctx& cx, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                    ConvGemmParams p, const Geo& g, long long images, cudaStream_t st,
                    const CUtensorMap* x0 = nullptr, const CUtensorMap* x1 = nullptr, const CUtensorMap* x2 = nullptr,
                    const CUtensorMap* x3 = nullptr, const char* name = "convgemm_kernel", const CUtensorMap* x4 = nullptr,
                    const CUtensorMap* x5 = nullptr, const CUtensorMap* x6 = nullptr) {
  p.B = static_cast<int>(images), p.H = g.H, p.W = g.W;
  p.BW = g.BW, p.BH = g.BH, p.tiles_w = g.tiles_w, p.tiles_h = g.tiles_h;
  p.num_m_tiles = static_cast<int>(images) * g.tiles_w * g.tiles_h;
  const DeviceInfo& dev = cx.dev;
  // staged epilogue: outputs go through swizzled shared memory + TMA stores (0: direct per-thread stores)
  p.staged = (x0 != nullptr && EPI != EPI_HEAD && cx.knobs.staged) ? 1 : 0;
  if (p.staged && EPI == EPI_LSTM && cx.knobs.staged == 2 && x4 != nullptr) p.staged = 2;
  {
    int kblocks = 0;
    for (int s = 0; s < p.nseg; ++s) kblocks += p.seg[s].chunks * p.seg[s].kh * p.seg[s].kw;
    p.rotate = (kblocks <= kKtabMax && kblocks > 1 && cx.knobs.rotate) ? 1 : 0;
  }
  const int stg_half = p.staged ? stg_half_bytes(EPI, p.staged) : 0;
  if (p.staged == 2) x0 = x4, x1 = x5, x2 = x6, x3 = nullptr;  // the 8-channel-box maps
  const int stage_bytes = kABytes + p.n_tile * 128;
  const int fixed = static_cast<int>(convgemm_smem_bytes(0, p.n_tile, p.n_tiles, stg_half));
  int stages = (dev.smem_optin - fixed) / stage_bytes;
  if (stages > cx.knobs.stages) stages = cx.knobs.stages;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return fail(CLSTM_EINVAL, "convgemm: not enough shared memory for n_tile=%d", p.n_tile);
  p.stages = stages;
  const size_t smem = convgemm_smem_bytes(stages, p.n_tile, p.n_tiles, stg_half);
  static bool attr_set = false;
  if (!attr_set) {
    CU_TRY(cudaFuncSetAttribute(convgemm_kernel<E, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                dev.smem_optin));
    attr_set = true;
  }
  const int total = p.num_m_tiles * p.n_tiles;
  const int grid = total < dev.sms ? total : dev.sms;
  convgemm_kernel<E, EPI><<<grid, kGemmThreads, smem, st>>>(a0, a1, b, x0 ? *x0 : b, x1 ? *x1 : b, x2 ? *x2 : b,
                                                            x3 ? *x3 : (x0 ? *x0 : b), p);
  return after_launch(name);
}

// Cell step on CTA pairs (cellstep_pair.cuh): shapes with at least two waves of tiles and an even number of tiles per
// image; `used` stays false otherwise (the caller then launches convgemm_kernel<E, EPI_LSTM>).
template <typename E>
int launch_cellstep_pair(const Ctx& cx, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b128,
                         ConvGemmParams p, const Geo& g, cudaStream_t st, const CUtensorMap& xc, const CUtensorMap& xh,
                         const CUtensorMap& xg, const char* name, bool* used) {
  *used = false;
  const DeviceInfo& dev = cx.dev;
  if (!cx.knobs.pair || cx.knobs.staged != 1 || p.n_tile != 256) return 0;
  const int tiles_img = g.tiles_w * g.tiles_h;
  const long long m_tiles = static_cast<long long>(g.B) * tiles_img;
  int kblocks = 0;
  for (int s = 0; s < p.nseg; ++s) kblocks += p.seg[s].chunks * p.seg[s].kh * p.seg[s].kw;
  if ((tiles_img & 1) || m_tiles * p.n_tiles < 2ll * dev.sms || kblocks > kKtabMax || kblocks < 1) return 0;
  p.B = g.B, p.H = g.H, p.W = g.W;
  p.BW = g.BW, p.BH = g.BH, p.tiles_w = g.tiles_w, p.tiles_h = g.tiles_h;
  p.num_m_tiles = static_cast<int>(m_tiles);
  p.staged = 1;
  p.rotate = (kblocks > 1 && cx.knobs.rotate) ? 1 : 0;
  int stages = (dev.smem_optin - static_cast<int>(cellstep_pair_smem_bytes(0, p.n_tiles))) / kPairStageBytes;
  if (stages > cx.knobs.stages) stages = cx.knobs.stages;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return 0;
  p.stages = stages;
  static bool attr_set = false;
  if (!attr_set) {
    CU_TRY(cudaFuncSetAttribute(cellstep_pair_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, dev.smem_optin));
    attr_set = true;
  }
  const int grid = dev.sms & ~1;
  cellstep_pair_kernel<E><<<grid, kGemmThreads, cellstep_pair_smem_bytes(stages, p.n_tiles), st>>>(a0, a1, b128, xc, xh, xg,
                                                                                                 xc, p);
  *used = true;
  return after_launch(name);
}

// Transposed dgrad (dgradT.cuh): channels as M, 256 pixels as N.
template <typename E>
int launch_dgradT(const Ctx& cx, const CUtensorMap& dz128, const CUtensorMap& wT, const CUtensorMap& x0,
                  const CUtensorMap& x1, const ConvSeg& seg, int rows_d, int split_col, const Geo& g, long long images,
                  cudaStream_t st, bool* used) {
  *used = false;
  const DeviceInfo& dev = cx.dev;
  if (g.BW < 16 || !cx.knobs.dgradT) return 0;
  DgradTParams p;
  memset(&p, 0, sizeof(p));
  p.B = static_cast<int>(images), p.H = g.H, p.W = g.W;
  p.BW = g.BW, p.BH = g.BH, p.tiles_w = g.tiles_w, p.tiles_h = g.tiles_h;
  p.num_m_tiles = static_cast<int>(images) * g.tiles_w * g.tiles_h;
  p.seg = seg;
  p.m_tiles = (rows_d + 127) / 128;
  p.split_col = split_col;
  p.lbw = 0;
  while ((1 << p.lbw) < g.BW) ++p.lbw;
  p.rotate = cx.knobs.rotate_d ? 1 : 0;  // measured: no gain for dgrad (weights are 1/3 of its operand bytes)
  int stages = (dev.smem_optin - static_cast<int>(dgradT_smem_bytes(0))) / kDtStageBytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return 0;
  p.stages = stages;
  static bool attr_set = false;
  if (!attr_set) {
    CU_TRY(cudaFuncSetAttribute(dgradT_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, dev.smem_optin));
    attr_set = true;
  }
  const int units = ((p.num_m_tiles + 1) / 2) * p.m_tiles;
  const int grid = units < dev.sms ? units : dev.sms;
  dgradT_kernel<E><<<grid, kGemmThreads, dgradT_smem_bytes(stages), st>>>(dz128, wT, x0, x1, p);
  *used = true;
  return after_launch("dgradT_kernel");
}

// The worker-warp generation of the fused kernel (dgradT_fused2.cuh) is selected and fits in shared memory.
inline bool fused2_ok(const Ctx& cx) {
  return cx.knobs.fuse_workers == 2 &&
         (cx.dev.smem_optin - static_cast<int>(dgradTf2_smem_bytes(0))) / kDtStageBytes >= 2;
}

// Transposed dgrad whose epilogue also runs the gate gradient of the NEXT cell step of the backward chain.
template <typename E>
int launch_dgradT_fused(const Ctx& cx, const CUtensorMap& dz128, const CUtensorMap& wT, const CUtensorMap& x0,
                        const CUtensorMap& x1, const ConvSeg& seg, const Geo& g, long long images, const GateFuse& f,
                        cudaStream_t st, const CUtensorMap* seg2_act = nullptr, const CUtensorMap* seg2_w = nullptr,
                        int seg2_kblocks = 0, bool state16 = false) {
  const DeviceInfo& dev = cx.dev;
  DgradTParams p;
  memset(&p, 0, sizeof(p));
  p.B = static_cast<int>(images), p.H = g.H, p.W = g.W;
  p.BW = g.BW, p.BH = g.BH, p.tiles_w = g.tiles_w, p.tiles_h = g.tiles_h;
  p.num_m_tiles = static_cast<int>(images) * g.tiles_w * g.tiles_h;
  p.seg = seg;
  p.m_tiles = 1;
  p.lbw = 0;
  while ((1 << p.lbw) < g.BW) ++p.lbw;
  p.rotate = cx.knobs.rotate_d ? 1 : 0;
  int stages = (dev.smem_optin - static_cast<int>(dgradTf_smem_bytes(0))) / kDtStageBytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return fail(CLSTM_EINVAL, "dgradT_fused: not enough shared memory");
  p.stages = stages;
  const int units = (p.num_m_tiles + 1) / 2;
  const int grid = units < dev.sms ? units : dev.sms;
  if (fused2_ok(cx) && f.src2 == nullptr && f.fuse_units >= units) {
    // second generation: drain warps + gate-gradient worker warps meeting at a staging ring (3 operand stages)
    int st2 = (dev.smem_optin - static_cast<int>(dgradTf2_smem_bytes(0))) / kDtStageBytes;
    if (st2 > kMaxStages) st2 = kMaxStages;
    if (st2 >= 2) {
      p.stages = st2;
      p.extra = seg2_kblocks;
      const CUtensorMap& gA = seg2_kblocks ? *seg2_act : dz128;
      const CUtensorMap& gW = seg2_kblocks ? *seg2_w : wT;
      static bool attr2_set = false;
      if (!attr2_set) {
        CU_TRY((cudaFuncSetAttribute(dgradT_fused2_kernel<E, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, dev.smem_optin)));
        CU_TRY((cudaFuncSetAttribute(dgradT_fused2_kernel<E, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dev.smem_optin)));
        CU_TRY((cudaFuncSetAttribute(dgradT_fused2_kernel<E, 2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dev.smem_optin)));
        attr2_set = true;
      }
      // c' recomputed from the saved gates: 11-bit (fp16) gates only, see gate_grad_item4
      const bool rc = cx.knobs.recomp_c && std::is_same<E, __half>::value;
      if (f.c16 && !rc) return fail(CLSTM_EINVAL, "dgradT_fused: a 16-bit c stack needs CLSTM_RECOMP_C=1");
      if (state16) {
        if (!rc)
          return fail(CLSTM_EINVAL, "dgradT_fused: 16-bit states need CLSTM_FUSE_WORKERS=2 and CLSTM_RECOMP_C=1");
        dgradT_fused2_kernel<E, 2, true, true><<<grid, kDf2Threads, dgradTf2_smem_bytes(st2), st>>>(dz128, wT, x1, gA, gW, p, f);
      } else if (rc)
        dgradT_fused2_kernel<E, 2, true><<<grid, kDf2Threads, dgradTf2_smem_bytes(st2), st>>>(dz128, wT, x1, gA, gW, p, f);
      else
        dgradT_fused2_kernel<E, 2><<<grid, kDf2Threads, dgradTf2_smem_bytes(st2), st>>>(dz128, wT, x1, gA, gW, p, f);
      return after_launch(seg2_kblocks ? "dgradT_fused2_kernel[+head dgrad]" : "dgradT_fused2_kernel");
    }
  }
  if (seg2_kblocks || state16 || f.c16)
    return fail(CLSTM_EINVAL, "dgradT_fused: the second K segment / 16-bit states need the worker-warp kernel");
  static bool attr_set = false;
  if (!attr_set) {
    CU_TRY(cudaFuncSetAttribute(dgradT_fused_kernel<E, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, dev.smem_optin));
    CU_TRY((cudaFuncSetAttribute(dgradT_fused_kernel<E, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, dev.smem_optin)));
    attr_set = true;
  }
  if (cx.knobs.fuse_sets == 4)
    dgradT_fused_kernel<E, 2, 4><<<grid, 128 + 2 * 128, dgradTf_smem_bytes(stages), st>>>(dz128, wT, x0, x1, p, f);
  else
    dgradT_fused_kernel<E, 2><<<grid, 128 + 2 * 128, dgradTf_smem_bytes(stages), st>>>(dz128, wT, x0, x1, p, f);
  return after_launch("dgradT_fused_kernel");
}

// Row-marching output head (head_rows.cuh).  `used` stays false when the shape is not supported (the caller then
// runs the implicit-GEMM head).
template <typename E>
int launch_head_rows(const Ctx& cx, const CUtensorMap& h128, const CUtensorMap& wz, HeadRowsParams p,
                     const Geo& g, cudaStream_t st, bool* used) {
  *used = false;
  const DeviceInfo& dev = cx.dev;
  if (g.BW != 128 || g.BH != 1 || p.c_out > 16 || !cx.knobs.head_rows) return 0;
  p.H = g.H, p.W = g.W;
  if (g.W <= 256) {
    p.strips = 1, p.strip_w = 256, p.x_halo = 0;
  } else {
    p.strip_w = 254, p.x_halo = 1, p.strips = (g.W + 253) / 254;
  }
  p.band_rows = cx.knobs.head_band > 0 ? cx.knobs.head_band : 32;
  p.bands = (g.H + p.band_rows - 1) / p.band_rows;
  int stages = (dev.smem_optin - static_cast<int>(head_rows_smem_bytes(p.chunks, 0))) / kABytes;
  if (stages > kHrMaxStages) stages = kHrMaxStages;
  if (stages < 3) return 0;
  p.stages = stages;
  const long long units = static_cast<long long>(p.images) * p.strips * p.bands;
  const int grid = units < dev.sms ? static_cast<int>(units) : dev.sms;
  const size_t smem = head_rows_smem_bytes(p.chunks, stages);
#define HR_LAUNCH(CO)                                                                                              \
  do {                                                                                                             \
    static bool attr_set = false;                                                                                  \
    if (!attr_set) {                                                                                               \
      CU_TRY(cudaFuncSetAttribute(head_rows_kernel<E, CO>, cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                                  dev.smem_optin));                                                                \
      attr_set = true;                                                                                             \
    }                                                                                                              \
    head_rows_kernel<E, CO><<<grid, kGemmThreads, smem, st>>>(h128, wz, p);                                        \
  } while (0)
  if (p.c_out <= 4)
    HR_LAUNCH(4);
  else if (p.c_out <= 8)
    HR_LAUNCH(8);
  else if (p.c_out <= 12)
    HR_LAUNCH(12);
  else
    HR_LAUNCH(16);
#undef HR_LAUNCH
  *used = true;
  return after_launch("head_rows_kernel");
}

template <typename E>
int launch_wgrad(const Ctx& cx, const CUtensorMap& a, const CUtensorMap& b0, const CUtensorMap& b1,
                 WgradParams p, const Geo& g, long long images, cudaStream_t st, const WgGateWork* gate = nullptr,
                 const char* name = "wgrad_kernel") {
  const DeviceInfo& dev = cx.dev;
  p.B = static_cast<int>(images), p.H = g.H, p.W = g.W;
  p.BW = g.BW2, p.BH = g.BH2, p.tiles_w = g.tiles_w2, p.tiles_h = g.tiles_h2;
  p.num_p_tiles = static_cast<int>(images) * g.tiles_w2 * g.tiles_h2;
  const int stage_bytes = p.halo ? (2 * kWgTileP * 128 + 2 * kWgHaloRowBytes) : (2 + p.group_size) * kWgTileP * 128;
  int stages = (dev.smem_optin - static_cast<int>(wgrad_smem_bytes(0, 0))) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return fail(CLSTM_EINVAL, "wgrad: not enough shared memory");
  p.stages = stages;
  const size_t smem = wgrad_smem_bytes(0, 0) + static_cast<size_t>(stages) * stage_bytes;
  static bool attr_set = false;
  if (!attr_set) {
    CU_TRY(cudaFuncSetAttribute(wgrad_kernel<E, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, dev.smem_optin));
    CU_TRY(cudaFuncSetAttribute(wgrad_kernel<E, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, dev.smem_optin));
    CU_TRY(cudaFuncSetAttribute(wgrad_kernel<E, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, dev.smem_optin));
    attr_set = true;
  }
  const int groups = (p.total_blocks + p.group_size - 1) / p.group_size;
  const int grid = groups * p.n_blocks * p.splits;
  if (gate != nullptr) {
    if (grid > kBiasRowsMax / 16) return fail(CLSTM_EINVAL, "wgrad + gate workers: grid %d too large", grid);
    if (cx.knobs.wg_gate == 2 || (cx.knobs.wg_gate == 0 && cx.knobs.hybrid_wg == 2))  // register-rebalanced workers
      wgrad_kernel<E, 2><<<grid, kWgThreads + kWgGateThreads2, smem, st>>>(a, b0, b1, p, *gate);
    else
      wgrad_kernel<E, 1><<<grid, kWgThreads + kWgGateThreads, smem, st>>>(a, b0, b1, p, *gate);
    return after_launch("wgrad_gate_kernel");
  }
  WgGateWork none;
  memset(&none, 0, sizeof(none));
  wgrad_kernel<E, 0><<<grid, kWgThreads, smem, st>>>(a, b0, b1, p, none);
  return after_launch(name);
}

// Row-tiled im2col when its shared-memory tile fits, else the generic gather kernels.
template <typename E, int MODE>
int launch_row_im2col(const float* s0, const float* s1, E* out, int B, int T, int C, int H, int W, int kh, int kw,
                      int KP, int t0, int nt, const float* scale_ptr, cudaStream_t st, bool* used) {
  const size_t smem = row_im2col_smem_bytes(C, W, kh, kw, KP);
  *used = false;
  if (smem > 160 * 1024 || static_cast<long long>(nt) * B * H > 0x7fffffffll) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    CU_TRY(cudaFuncSetAttribute(row_im2col_kernel<E, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set = true;
  }
  row_im2col_kernel<E, MODE><<<nt * B * H, 256, smem, st>>>(s0, s1, out, B, T, C, H, W, kh, kw, KP, t0, scale_ptr);
  *used = true;
  return after_launch(MODE == 1 ? "head_dlogit_im2col" : "x_im2col");
}

// --------------------------------------------------------------------------- cell building blocks
// Repack one cell's reference-layout parameters (layers/ConvLSTM.py:34-40 weight/bias).
template <typename E>
int pack_cell(const Ctx& ctx, CellState& cs, const float* w, const float* bias, cudaStream_t st) {
  pack_cell_weights_fwd_kernel<E><<<kPackBlocks, 256, 0, st>>>(w, bias, static_cast<E*>(cs.wp), cs.bias_p, cs.g);
  RC_TRY(after_launch("pack_cell_weights_fwd_kernel"));
  if (ctx.training) {
    pack_cell_weights_dgrad_kernel<E><<<kPackBlocks, 256, 0, st>>>(w, static_cast<E*>(cs.wd), cs.g, cs.with_x);
    RC_TRY(after_launch("pack_cell_weights_dgrad_kernel"));
  }
  return 0;
}

// One fused cell step (layers/ConvLSTM.py:42-57): reads the input tensor and h slot `sp`, c_prev;
// writes h slot `sn`, c_next (and the gates when training).
template <typename E>
int cell_forward_step(const Ctx& ctx, CellState& cs, const InputRef& in, int sp, int sn, const float* c_prev,
                      float* c_next, void* gates, cudaStream_t st, int cprev_slot = -1, int cnext_slot = -1,
                      int gates_step = -1, bool zero_h = false) {
  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  const CellGeom& g = cs.g;
  p.n_tiles = ctx.HP / 64;
  p.n_tile = 256;
  p.nseg = zero_h ? 1 : 2;  // h == 0 (a cell's first step of a rollout): the h segment of the K loop is skipped
  if (g.in_col)
    p.seg[0] = ConvSeg{g.KIN / 64, 1, 1, in.b_off};
  else
    p.seg[0] = ConvSeg{g.CIP / 64, g.kh, g.kw, in.b_off};
  p.seg[1] = ConvSeg{ctx.HP / 64, g.kh, g.kw, sp * ctx.geo.B};
  p.bias = cs.bias_p;
  p.c_prev = c_prev;
  p.c_next = c_next;
  p.h_next = static_cast<E*>(cs.h) + static_cast<size_t>(sn) * cs.h_slot_elems(ctx.geo);
  p.gates = gates;
  p.ldc = ctx.HP;
  p.c16 = ctx.c16 ? 1 : 0;  // only plans whose every cell step takes the staged epilogue below set it (decide_c16)
  if (cnext_slot >= 0) {  // staged epilogue: image offsets into the c / h / gates stacks
    p.cprev_boff = (c_prev != nullptr && cprev_slot >= 0) ? cprev_slot * ctx.geo.B : -1;
    p.cnext_boff = cnext_slot * ctx.geo.B;
    p.hnext_boff = sn * ctx.geo.B;
    p.gates_boff = (gates != nullptr && gates_step >= 0) ? gates_step * ctx.geo.B : -1;
    if (!g.in_col || ctx.knobs.pair >= 2) {  // the short-K bottom cell is epilogue bound: measured 2 % slower on pairs
      bool used = false;
      RC_TRY((launch_cellstep_pair<E>(ctx, *in.map128, cs.m_h128, cs.m_wp128, p, ctx.geo, st, cs.m_c16, cs.m_h16,
                                      ctx.training ? cs.m_g16 : cs.m_h16, g.in_col ? "cell_step[x im2col, pair]" : "cell_step[pair]",
                                      &used)));
      if (used) return 0;
    }
    return launch_convgemm<E, EPI_LSTM>(ctx, *in.map128, cs.m_h128, cs.m_wp, p, ctx.geo, ctx.geo.B, st, &cs.m_c16,
                                        &cs.m_h16, ctx.training ? &cs.m_g16 : &cs.m_h16, nullptr,
                                        g.in_col ? "cell_step[x im2col]" : "cell_step", &cs.m_c8, &cs.m_h8,
                                        ctx.training ? &cs.m_g8 : &cs.m_h8);
  }
  return launch_convgemm<E, EPI_LSTM>(ctx, *in.map128, cs.m_h128, cs.m_wp, p, ctx.geo, ctx.geo.B, st, nullptr, nullptr,
                                      nullptr, nullptr, "cell_step[direct stores]");
}

// Backward of one cell step, three launches: fused gate gradient -> dgrad (dx | dh_prev) -> wgrad
// accumulation.  dh sources (fp32 NHWC, scaled by S) may be null.  c_prev may be null (zeros).
template <typename E>
int cell_gate_grad(const Ctx& ctx, CellState& cs, const void* gates, const float* c_prev, const float* c_next,
                   const float* dh0, const float* dh1, const float* dh2, int first, cudaStream_t st, int buf = 0,
                   bool state16 = false) {
  // pointwise gate gradient (backward of layers/ConvLSTM.py:48-55).  (Forcing the max-shared-memory carveout so
  // it could co-reside with a wgrad CTA slows it from 442 to 614 us — it needs L1 for its loads in flight — and the
  // two kernels still did not overlap: DESIGN.md "backward overlap".)
  gate_grad_kernel<E><<<kGateGradBlocks, 256, 256 * 9 * sizeof(float), st>>>(
      static_cast<const E*>(gates), c_prev, c_next, dh0, dh1, dh2, cs.dc, static_cast<E*>(ctx.dzb[buf]), cs.bpart,
      !first, ctx.geo.npix(), ctx.HP, ctx.amax + 1, state16 ? 1 : 0, ctx.c16 ? 1 : 0);
  return after_launch("gate_grad_kernel");
}

template <typename E>
int cell_dgrad(const Ctx& ctx, CellState& cs, cudaStream_t st, int buf = 0) {
  // d[x | h_prev] = conv_transpose(dz, W)  (backward of layers/ConvLSTM.py:45-47)
  const CellGeom& g = cs.g;
  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  p.n_tile = cs.n_tile_d;
  p.n_tiles = cs.rows_d / cs.n_tile_d;
  p.nseg = 1;
  p.seg[0] = ConvSeg{4 * ctx.HP / 64, g.kh, g.kw, 0};
  p.out0 = cs.dxb;
  p.out1 = cs.dh_own;
  p.split_col = cs.with_x ? g.CIP : 0;
  p.ld0 = g.CIP;
  p.ld1 = ctx.HP;
  p.out_scale = 1.f;
  {
    bool used = false;
    RC_TRY((launch_dgradT<E>(ctx, ctx.m_dz128b[buf], cs.m_wdT, cs.with_x ? cs.m_dxT : cs.m_dhT, cs.m_dhT, p.seg[0],
                             cs.rows_d, p.split_col, ctx.geo, ctx.geo.B, st, &used)));
    if (used) return 0;
  }
  return launch_convgemm<E, EPI_STORE>(ctx, ctx.m_dz128b[buf], ctx.m_dz128b[buf], cs.m_wd, p, ctx.geo, ctx.geo.B, st,
                                       cs.with_x ? &cs.m_dx16 : &cs.m_dh16, &cs.m_dh16, &cs.m_dh16, nullptr,
                                       "dgrad[pixel-major]");
}

// slot of a cell's h / c state after `s` steps (s = 0: initial zeros)
inline int hslot(const CellState& cs, int s) { return s % cs.slots_h; }
inline int cslot(const CellState& cs, int s) { return s % cs.slots_c; }
// Address of slot `slot` of a cell's c stack: fp32 [npix][HP], or E [npix][HP] times kCScale when ctx.c16 (the float-typed
// pointer is only a handle for the kernels, which know the format).
inline float* cptr(const Ctx& ctx, const CellState& cs, int slot) {
  const size_t slot_floats = ctx.geo.npix() * ctx.HP / (ctx.c16 ? 2 : 1);
  return cs.c + static_cast<size_t>(slot) * slot_floats;
}

// dgrad of `cs` (reading dz[buf]) with the gate gradient of the NEXT step of the backward chain — cell `cn`, time
// step `nt`, dh sources own / e1 / e2 — fused into its epilogue (dgradT.cuh); that gate gradient lands in dz[buf^1].
// A source equal to cs.dxb is the dx this launch produces: it is consumed from shared memory and never written.
template <typename E>
int cell_dgrad_fused(const Ctx& ctx, CellState& cs, CellState& cn, int nt, const float* own, const float* e1,
                     const float* e2, int buf, cudaStream_t st, int fuse_units = 0x7fffffff,
                     const CUtensorMap* seg2_act = nullptr, const CUtensorMap* seg2_w = nullptr, int seg2_kblocks = 0,
                     bool state16 = false) {
  const size_t npix = ctx.geo.npix();
  const int HP = ctx.HP;
  GateFuse f;
  memset(&f, 0, sizeof(f));
  f.gates = static_cast<const E*>(cn.gates) + static_cast<size_t>(nt) * npix * 4 * HP;
  f.c_prev = (nt == 0) ? nullptr : cptr(ctx, cn, cslot(cn, nt));
  f.c_next = cptr(ctx, cn, cslot(cn, nt + 1));
  f.c16 = ctx.c16 ? 1 : 0;
  const float* srcs[3] = {own, e1, e2};
  for (const float*& sp : srcs)
    if (sp != nullptr && cs.with_x && sp == cs.dxb) {
      f.use_stg = 1;
      sp = nullptr;
    }
  f.src0 = srcs[0], f.src1 = srcs[1], f.src2 = srcs[2];
  f.h_block = cs.with_x ? 1 : 0;
  f.dc = cn.dc;
  f.dz_out = ctx.dzb[buf ^ 1];
  f.bias_partial = cn.bpart;
  f.dz_absmax = ctx.amax + 1;
  f.HP = HP;
  f.fuse_units = fuse_units;
  f.pf_dist = ctx.knobs.fuse_pf;
  return launch_dgradT_fused<E>(ctx, ctx.m_dz128b[buf], cs.m_wdT, cs.with_x ? cs.m_dxT : cs.m_dhT,
                                state16 ? cs.m_dhT16 : cs.m_dhT, ConvSeg{4 * HP / 64, cs.g.kh, cs.g.kw, 0}, ctx.geo,
                                ctx.geo.B, f, st, seg2_act, seg2_w, seg2_kblocks, state16);
}

// The 16-bit recurrent gradient states (CLSTM_STATE16) apply when every dgrad of the backward chain is the worker-warp
// fused kernel with recomputed c' (see plan_backward): knobs and dtype part of the condition.
template <typename E>
inline bool state16_knobs_ok(const Ctx& ctx) {
  return ctx.knobs.state16 && ctx.knobs.recomp_c && ctx.knobs.fuse_workers == 2 && fused2_ok(ctx) &&
         std::is_same<E, __half>::value;
}

// Shapes the fused dgrad + gate-gradient kernel supports (hidden padded to 64, 32-bit element offsets).
inline bool fuse_supported(const Ctx& ctx) {
  return ctx.HP == 64 && ctx.geo.BW >= 16 && ctx.geo.npix() * 4 * ctx.HP < (1ull << 32) && ctx.knobs.dgradT &&
         ctx.knobs.fuse_gate;
}

template <typename E>
int cell_wgrad(const Ctx& ctx, CellState& cs, const InputRef& in, int sp, int first, cudaStream_t st, int buf = 0,
               const WgGateWork* gate = nullptr) {
  // dW += im2col([x, h_prev])^T dz, accumulated over the cell's time steps
  const CellGeom& g = cs.g;
  const Geo& geo = ctx.geo;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.n_blocks = 4 * ctx.HP / 128;
  p.group_size = cs.wg_group;
  p.total_blocks = cs.wg_total;
  if (g.in_col)
    p.seg[0] = WgradSeg{g.KIN / 64, g.KIN / 64, 1, 0, 0, in.b_off};
  else
    p.seg[0] = WgradSeg{g.kh * g.kw * (g.CIP / 64), g.CIP / 64, g.kw, g.kh / 2, g.kw / 2, in.b_off};
  p.seg[1] = WgradSeg{g.kh * g.kw * (ctx.HP / 64), ctx.HP / 64, g.kw, g.kh / 2, g.kw / 2, sp * geo.B};
  p.a_b_off = 0;
  p.splits = cs.wg_splits;
  p.partial = cs.wpart;
  p.accumulate = !first;
  // halo rows: both segments 3x3 convs over one 64-channel chunk, 64-pixel single-row K steps, groups of 6 taps
  if (!g.in_col && g.kh == 3 && g.kw == 3 && g.CIP == 64 && ctx.HP == 64 && geo.BW2 == 64 && geo.BH2 == 1 &&
      cs.wg_group == 6 && cs.wg_total == 18 && in.map66 != nullptr && ctx.knobs.wg_halo) {
    p.halo = 1;
    return launch_wgrad<E>(ctx, ctx.m_dz64b[buf], *in.map66, cs.m_h66, p, geo, geo.B, st, gate, "wgrad[halo rows]");
  }
  return launch_wgrad<E>(ctx, ctx.m_dz64b[buf], *in.map64, cs.m_h64, p, geo, geo.B, st, gate,
                         g.in_col ? "wgrad[x im2col]" : "wgrad[boxes]");
}

template <typename E>
int cell_backward_step(const Ctx& ctx, CellState& cs, const InputRef& in, int sp, const void* gates,
                       const float* c_prev, const float* c_next, const float* dh0, const float* dh1,
                       const float* dh2, cudaStream_t st, bool need_dgrad = true) {
  const int first = cs.bwd_started ? 0 : 1;
  cs.bwd_started = true;
  RC_TRY(cell_gate_grad<E>(ctx, cs, gates, c_prev, c_next, dh0, dh1, dh2, first, st));
  if (need_dgrad) RC_TRY(cell_dgrad<E>(ctx, cs, st));
  RC_TRY(cell_wgrad<E>(ctx, cs, in, sp, first, st));
  return 0;
}

// Reduce the split partial sums into reference-layout gradients (unscaled by 1/S).
int cell_finalize(const Ctx& ctx, CellState& cs, float* dw, float* db, int accumulate, cudaStream_t st,
                  int bias_rows = kGateGradBlocks) {
  const CellGeom& g = cs.g;
  if (dw) {
    cell_wgrad_finalize_kernel<<<kPackBlocks, 256, 0, st>>>(cs.wpart, dw, g, cs.wg_splits, ctx.scale + 1,
                                                            accumulate);
    RC_TRY(after_launch("cell_wgrad_finalize_kernel"));
  }
  if (db) {
    cell_bias_finalize_kernel<<<(4 * g.hid + 31) / 32, 256, 0, st>>>(cs.bpart, db, bias_rows, g.hid,
                                                                     ctx.HP, ctx.scale + 1, accumulate);
    RC_TRY(after_launch("cell_bias_finalize_kernel"));
  }
  return 0;
}

int validate_common(int batch, int height, int width, int in_channels, int hidden, int kh, int kw, int dtype) {
  if (batch < 1 || height < 1 || width < 1 || in_channels < 1 || hidden < 1)
    return fail(CLSTM_EINVAL, "batch/height/width/channels/hidden must be positive");
  if (kh < 1 || kw < 1 || (kh % 2) == 0 || (kw % 2) == 0 || kh > 7 || kw > 7)
    return fail(CLSTM_EINVAL,
                "kernel size (%d,%d) unsupported: odd sizes up to 7 only (the reference itself breaks on even "
                "kernels, layers/ConvLSTM.py:32-54)",
                kh, kw);
  if (pad_hidden(hidden) < 0) return fail(CLSTM_EINVAL, "hidden_dim %d > 512 unsupported", hidden);
  if (dtype != CLSTM_F16 && dtype != CLSTM_BF16) return fail(CLSTM_EINVAL, "dtype must be CLSTM_F16 or CLSTM_BF16");
  if (static_cast<long long>(batch) * height * width > (1ll << 30))
    return fail(CLSTM_EINVAL, "batch*height*width too large");
  return 0;
}

#define DISPATCH_E(dtype, call)          \
  ((dtype) == CLSTM_F16 ? call(__half) : call(__nv_bfloat16))

}  // namespace

// ============================================================================ rollout plan
struct clstm_plan {
  clstm_config_t cfg;
  Ctx ctx;
  int L = 0, ncell = 0, KX = 0, KG = 0, NT = 0;
  size_t ws_bytes = 0;
  bool bound = false, weights_set = false, forward_done = false;
  std::vector<CellState> cells;
  void* xcol = nullptr;       // E [T_in*B][H][W][KX]
  void* G = nullptr;          // E [npix][KG]
  float* dstack = nullptr;    // fp32 [npix][HP]
  void* wh = nullptr;         // E [NT][9HP]
  void* wz = nullptr;         // E [HP/64][144][64]: taps as output columns (head_rows.cuh); null if c_out > 16
  float* bias_h = nullptr;    // fp32 [NT]
  void* whd = nullptr;        // E [n_tile_hd rows = HP][KG]
  void* whdT = nullptr;       // E [128][KG]: rows [0, HP) = whd, the rest zero — the A operand of the head's dgrad when it
                              // rides inside the fused dgrad of decoder cell 0 (rows = that cell's x | h channel split)
  float* hpart = nullptr;     // fp32 [splits][128*(KG/128)][HP]
  float* hbpart = nullptr;    // fp32 [C_out][batch * hb_chunks]
  int hb_chunks = 1;          // pieces per (b, c) run in head_grad_stats_kernel
  int head_splits = 0, head_group = 0, n_tile_hd = 0;
  CUtensorMap m_xcol128, m_xcol64, m_G128, m_G64, m_wh, m_whd, m_whdT, m_dstack16, m_wz;
  // persistent forward chain (rollout_persist.cuh): device copies of the tensor maps and of the step table
  bool persist_ok = false;
  int persist_stages = 0;
  void* pmaps = nullptr;
  void* psteps = nullptr;
  unsigned int* pcounter = nullptr;
  PersistParams pparams;
  // CUDA-graph replay of the forward / backward launch sequences (launch-bound shapes only)
  bool graph_ok = false;
  GraphCache g_fwd, g_bwd;
};

namespace {

void carve_plan(clstm_plan* p, uint8_t* base) {
  Carver cv;
  cv.base = base;
  const clstm_config_t& c = p->cfg;
  Ctx& ctx = p->ctx;
  const size_t npix = ctx.geo.npix();
  const int HP = ctx.HP;
  ctx.scale = cv.take<float>(64);
  ctx.amax = reinterpret_cast<unsigned int*>(cv.take<float>(64));
  p->xcol = cv.take<void>(static_cast<size_t>(c.t_in) * npix * p->KX * 2);
  for (int k = 0; k < p->ncell; ++k) carve_cell(cv, cv, p->cells[k], ctx);
  p->wh = cv.take<void>(static_cast<size_t>(p->NT) * 9 * HP * 2);
  p->bias_h = cv.take<float>(static_cast<size_t>(p->NT) * 4);
  p->wz = (c.out_channels <= 16) ? cv.take<void>(static_cast<size_t>(HP / 64) * kHrN * 64 * 2) : nullptr;
  p->pmaps = cv.take<void>(static_cast<size_t>(kPersistMaxMaps) * sizeof(CUtensorMap));
  p->psteps = cv.take<void>(static_cast<size_t>(p->L) * (c.t_in + c.t_out) * sizeof(PersistStep));
  p->pcounter = reinterpret_cast<unsigned int*>(cv.take<float>(64));
  if (c.training) {
    ctx.dzb[0] = cv.take<void>(npix * 4 * HP * 2);
    ctx.dzb[1] = cv.take<void>(npix * 4 * HP * 2);
    ctx.dz = ctx.dzb[0];
    p->G = cv.take<void>(npix * p->KG * 2);
    p->dstack = cv.take<float>(npix * HP * 4);
    p->whd = cv.take<void>(static_cast<size_t>(HP) * p->KG * 2);
    p->whdT = cv.take<void>(static_cast<size_t>(128) * p->KG * 2);
    p->hpart = cv.take<float>(static_cast<size_t>(p->head_splits) * p->KG * HP * 4);
    {  // about 8 blocks per SM, at least 4096 elements per block
      const size_t per_b = static_cast<size_t>(c.t_out) * c.height * c.width;
      int chunks = (8 * 148 + c.batch * c.out_channels - 1) / (c.batch * c.out_channels);
      const size_t max_chunks = per_b / 4096 > 0 ? per_b / 4096 : 1;
      if (static_cast<size_t>(chunks) > max_chunks) chunks = static_cast<int>(max_chunks);
      p->hb_chunks = chunks < 1 ? 1 : chunks;
    }
    p->hbpart = cv.take<float>(static_cast<size_t>(c.out_channels) * c.batch * p->hb_chunks * 4);
  }
  p->ws_bytes = align_up(cv.off, 1024);
}

// Input of cell k at its step t (conv_lstm.py:176-196).
InputRef plan_input(clstm_plan* p, int k, int t) {
  InputRef in;
  const int B = p->cfg.batch;
  const int L = p->L;
  if (k == 0) {
    in.map128 = &p->m_xcol128, in.map64 = &p->m_xcol64, in.b_off = t * B;  // x[:, t] (:177)
  } else if (k == L) {
    // decoder_1 input: encoder_vector (:185, :189) = last encoder h at t == 0, else last decoder h (:195)
    const CellState& src = (t == 0) ? p->cells[L - 1] : p->cells[p->ncell - 1];
    const int s = (t == 0) ? hslot(src, p->cfg.t_in) : hslot(src, t);
    in.map128 = &src.m_h128, in.map64 = &src.m_h64, in.map66 = &src.m_h66, in.b_off = s * B;
  } else {
    const CellState& src = p->cells[k - 1];  // the layer below, already stepped to t + 1 (:180, :192)
    in.map128 = &src.m_h128, in.map64 = &src.m_h64, in.map66 = &src.m_h66, in.b_off = hslot(src, t + 1) * B;
  }
  return in;
}

// Persistent forward chain (rollout_persist.cuh): eligible when every (pixel tile, 64-channel slice) gets its own CTA
// and the cell states fit in TMEM next to the accumulator.  Builds the device-side tensor-map table and step table.
int persist_setup(clstm_plan* p, cudaStream_t st) {
  const clstm_config_t& c = p->cfg;
  Ctx& ctx = p->ctx;
  const Geo& geo = ctx.geo;
  p->persist_ok = false;
  const int n_tiles = ctx.HP / 64;
  const long long tiles = static_cast<long long>(geo.B) * geo.tiles_w * geo.tiles_h * n_tiles;
  if (!ctx.knobs.persist || !ctx.knobs.staged || p->ncell > kPersistMaxCells || tiles > ctx.dev.sms || ctx.c16) return 0;
  if (1 + 5 * p->ncell > kPersistMaxMaps) return 0;
  for (int k = 0; k < p->ncell; ++k)
    if (p->cells[k].Kf / 64 > kKtabMax) return 0;
  int stages = (ctx.dev.smem_optin - static_cast<int>(persist_smem_bytes(0, p->ncell))) / (kABytes + 256 * 128);
  if (stages > ctx.knobs.stages) stages = ctx.knobs.stages;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return 0;
  p->persist_stages = stages;
  std::vector<CUtensorMap> maps(kPersistMaxMaps);
  memset(maps.data(), 0, maps.size() * sizeof(CUtensorMap));
  maps[0] = p->m_xcol128;
  PersistParams& pp = p->pparams;
  memset(&pp, 0, sizeof(pp));
  for (int k = 0; k < p->ncell; ++k) {
    const CellState& cs = p->cells[k];
    const int base = 1 + 5 * k;
    maps[base + 0] = cs.m_h128, maps[base + 1] = cs.m_wp, maps[base + 2] = cs.m_c16, maps[base + 3] = cs.m_h16;
    maps[base + 4] = c.training ? cs.m_g16 : cs.m_h16;
    PersistCell& pc = pp.cells[k];
    const CellGeom& g = cs.g;
    pc.seg[0] = g.in_col ? ConvSeg{g.KIN / 64, 1, 1, 0} : ConvSeg{g.CIP / 64, g.kh, g.kw, 0};
    pc.seg[1] = ConvSeg{ctx.HP / 64, g.kh, g.kw, 0};
    pc.kblocks = cs.Kf / 64;
    pc.kb_first = pc.seg[0].chunks * pc.seg[0].kh * pc.seg[0].kw;
    pc.map_a1 = base + 0, pc.map_b = base + 1, pc.map_xc = base + 2, pc.map_xh = base + 3, pc.map_xg = base + 4;
    pc.bias = cs.bias_p;
  }
  std::vector<PersistStep> steps;
  const int L = p->L, B = c.batch;
  auto add = [&](int k, int t) {
    const CellState& cs = p->cells[k];
    PersistStep s;
    memset(&s, 0, sizeof(s));
    s.cell = k;
    if (k == 0) {
      s.map_a0 = 0, s.a0_boff = t * B;  // x[:, t] (conv_lstm.py:177)
    } else if (k == L) {
      const int src = (t == 0) ? L - 1 : p->ncell - 1;  // encoder_vector (:185) / last decoder h (:195)
      s.map_a0 = 1 + 5 * src;
      s.a0_boff = ((t == 0) ? hslot(p->cells[src], c.t_in) : hslot(p->cells[src], t)) * B;
    } else {
      s.map_a0 = 1 + 5 * (k - 1);
      s.a0_boff = hslot(p->cells[k - 1], t + 1) * B;
    }
    s.a1_boff = hslot(cs, t) * B;
    s.hnext_boff = hslot(cs, t + 1) * B;
    s.cnext_boff = cslot(cs, t + 1) * B;
    s.gates_boff = t * B;
    s.first = (t == 0) ? 1 : 0;
    s.store_c = (c.training || t == cs.T - 1) ? 1 : 0;
    steps.push_back(s);
  };
  for (int t = 0; t < c.t_in; ++t)
    for (int l = 0; l < L; ++l) add(l, t);
  for (int t = 0; t < c.t_out; ++t)
    for (int l = 0; l < L; ++l) add(L + l, t);
  CU_TRY(cudaMemcpyAsync(p->pmaps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, st));
  CU_TRY(cudaMemcpyAsync(p->psteps, steps.data(), steps.size() * sizeof(PersistStep), cudaMemcpyHostToDevice, st));
  CU_TRY(cudaStreamSynchronize(st));  // the host vectors die with this scope
  pp.B = geo.B, pp.H = geo.H, pp.W = geo.W, pp.BW = geo.BW, pp.BH = geo.BH, pp.tiles_w = geo.tiles_w, pp.tiles_h = geo.tiles_h;
  pp.num_m_tiles = geo.B * geo.tiles_w * geo.tiles_h;
  pp.n_tiles = n_tiles;
  pp.ldc = ctx.HP;
  pp.stages = stages, pp.nsteps = static_cast<int>(steps.size()), pp.ncell = p->ncell, pp.training = c.training ? 1 : 0;
  pp.rotate = ctx.knobs.rotate ? 1 : 0;
  pp.maps = static_cast<const CUtensorMap*>(p->pmaps);
  pp.steps = static_cast<const PersistStep*>(p->psteps);
  pp.counter = p->pcounter;
  p->persist_ok = true;
  return 0;
}

template <typename E>
int launch_persist(clstm_plan* p, cudaStream_t st) {
  const size_t smem = persist_smem_bytes(p->persist_stages, p->ncell);
  static bool attr_set = false;
  if (!attr_set) {
    CU_TRY(cudaFuncSetAttribute(rollout_persist_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                p->ctx.dev.smem_optin));
    attr_set = true;
  }
  CU_TRY(cudaMemsetAsync(p->pcounter, 0, 4, st));
  void* args[] = {&p->pparams};
  const int grid = p->pparams.num_m_tiles * p->pparams.n_tiles;
  // cooperative launch: every CTA must be resident, each one waits for the h tiles of all others between steps
  CU_TRY(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(rollout_persist_kernel<E>), dim3(grid), dim3(kGemmThreads),
                                     args, smem, st));
  return after_launch("rollout_persist_kernel");
}

template <typename E>
int plan_set_weights(clstm_plan* p, const float* const* params, cudaStream_t st) {
  for (int k = 0; k < p->ncell; ++k) RC_TRY(pack_cell<E>(p->ctx, p->cells[k], params[2 * k], params[2 * k + 1], st));
  const float* wh = params[2 * p->ncell];
  const float* bh = params[2 * p->ncell + 1];
  pack_head_weights_fwd_kernel<E><<<kPackBlocks, 256, 0, st>>>(wh, bh, static_cast<E*>(p->wh), p->bias_h,
                                                               p->cfg.out_channels, p->cfg.hidden, p->ctx.HP, p->NT);
  RC_TRY(after_launch("pack_head_weights_fwd_kernel"));
  if (p->wz) {
    pack_head_weights_rows_kernel<E><<<kPackBlocks, 256, 0, st>>>(wh, static_cast<E*>(p->wz), p->cfg.out_channels,
                                                                p->cfg.hidden, p->ctx.HP);
    RC_TRY(after_launch("pack_head_weights_rows_kernel"));
  }
  if (p->cfg.training) {
    pack_head_weights_dgrad_kernel<E><<<kPackBlocks, 256, 0, st>>>(wh, static_cast<E*>(p->whd), p->cfg.out_channels,
                                                                   p->cfg.hidden, p->ctx.HP, p->KG);
    RC_TRY(after_launch("pack_head_weights_dgrad_kernel"));
    if (p->ctx.HP <= 128) {
      CU_TRY(cudaMemsetAsync(p->whdT, 0, static_cast<size_t>(128) * p->KG * 2, st));
      pack_head_weights_dgrad_kernel<E><<<kPackBlocks, 256, 0, st>>>(wh, static_cast<E*>(p->whdT), p->cfg.out_channels,
                                                                     p->cfg.hidden, p->ctx.HP, p->KG);
      RC_TRY(after_launch("pack_head_weights_dgrad_kernel"));
    }
  }
  return 0;
}

template <typename E>
int plan_forward(clstm_plan* p, const float* x, float* y, cudaStream_t st, int channels_last = 0) {
  const clstm_config_t& c = p->cfg;
  const Ctx& ctx = p->ctx;
  const Geo& geo = ctx.geo;
  const size_t npix = geo.npix();
  const int HP = ctx.HP;
  const int L = p->L;
  // x (B,T,C,H,W) -> im2col'd 16-bit tensor, once for all T_in steps
  bool tiled = false;
  if (channels_last)
    RC_TRY((launch_row_im2col<E, 2>(x, nullptr, static_cast<E*>(p->xcol), c.batch, c.t_in, c.in_channels, c.height,
                                    c.width, c.kernel_h, c.kernel_w, p->KX, 0, c.t_in, nullptr, st, &tiled)));
  else
    RC_TRY((launch_row_im2col<E, 0>(x, nullptr, static_cast<E*>(p->xcol), c.batch, c.t_in, c.in_channels, c.height,
                                    c.width, c.kernel_h, c.kernel_w, p->KX, 0, c.t_in, nullptr, st, &tiled)));
  if (!tiled) {
    pack_xcol_kernel<E><<<kPackBlocks, 256, 0, st>>>(x, static_cast<E*>(p->xcol), c.batch, c.t_in, c.in_channels,
                                                     c.height, c.width, c.kernel_h, c.kernel_w, p->KX, channels_last);
    RC_TRY(after_launch("pack_xcol_kernel"));
  }
  if (!c.training) {
    // ring-buffered states: slot 0 must read as the zero initial state (layers/ConvLSTM.py:59-64)
    for (int k = 0; k < p->ncell; ++k) {
      CellState& cs = p->cells[k];
      CU_TRY(cudaMemsetAsync(static_cast<uint8_t*>(cs.h) + static_cast<size_t>(hslot(cs, 0)) * npix * HP * 2, 0,
                             npix * HP * 2, st));
    }
  }
  auto step = [&](int k, int t) -> int {
    CellState& cs = p->cells[k];
    const InputRef in = plan_input(p, k, t);
    const float* c_prev = (t == 0) ? nullptr : cptr(ctx, cs, cslot(cs, t));
    float* c_next = cptr(ctx, cs, cslot(cs, t + 1));
    void* gates = c.training ? static_cast<void*>(static_cast<E*>(cs.gates) + static_cast<size_t>(t) * npix * 4 * HP)
                             : nullptr;
    return cell_forward_step<E>(ctx, cs, in, hslot(cs, t), hslot(cs, t + 1), c_prev, c_next, gates, st, cslot(cs, t),
                                cslot(cs, t + 1), t, /*zero_h=*/t == 0);
  };
  if (p->persist_ok) {
    RC_TRY(launch_persist<E>(p, st));  // the whole chain below in one launch, c resident in TMEM
  } else {
    for (int t = 0; t < c.t_in; ++t)  // conv_lstm.py:176-183
      for (int l = 0; l < L; ++l) RC_TRY(step(l, t));
    for (int t = 0; t < c.t_out; ++t)  // conv_lstm.py:188-196
      for (int l = 0; l < L; ++l) RC_TRY(step(L + l, t));
  }
  // head: Conv3d(1,3,3) + Sigmoid over the last-decoder h of every output step (conv_lstm.py:198-201).  The
  // h slots are permuted in memory, so the head runs once per output frame (B images each) and writes frame t
  // of y directly — the reference's stack / permute copies (:198-199) never exist.
  {
    const CellState& last = p->cells[p->ncell - 1];
    const bool contiguous = true;  // slots 1..T_out of the last decoder's h stack are adjacent in memory
    const int launches = 1;
    for (int t = 0; t < launches; ++t) {
      ConvGemmParams hp;
      memset(&hp, 0, sizeof(hp));
      hp.n_tile = p->NT;
      hp.n_tiles = 1;
      hp.nseg = 1;
      hp.seg[0] = ConvSeg{HP / 64, 3, 3, hslot(last, t + 1) * c.batch};
      hp.bias = p->bias_h;
      hp.y = y;
      hp.c_out = c.out_channels;
      hp.t_out = c.t_out;
      hp.b_img = c.batch;
      hp.t0 = t;
      const long long images = contiguous ? static_cast<long long>(c.t_out) * c.batch : c.batch;
      if (p->wz) {
        HeadRowsParams rp;
        memset(&rp, 0, sizeof(rp));
        rp.images = static_cast<int>(images);
        rp.img_off = hp.seg[0].b_off;
        rp.b_img = c.batch, rp.t0 = t, rp.t_out = c.t_out, rp.c_out = c.out_channels;
        rp.chunks = HP / 64;
        rp.bias = p->bias_h;
        rp.y = y;
        bool used = false;
        RC_TRY((launch_head_rows<E>(ctx, last.m_h128, p->m_wz, rp, geo, st, &used)));
        if (used) continue;
      }
      RC_TRY((launch_convgemm<E, EPI_HEAD>(ctx, last.m_h128, last.m_h128, p->m_wh, hp, geo, images, st, nullptr, nullptr,
                                           nullptr, nullptr, "head_conv[implicit GEMM]")));
    }
  }
  p->forward_done = true;
  return 0;
}

template <typename E>
int plan_backward(clstm_plan* p, const float* dy, const float* y, float* const* grads, int accumulate,
                  cudaStream_t st) {
  const clstm_config_t& c = p->cfg;
  Ctx& ctx = p->ctx;
  const Geo& geo = ctx.geo;
  const size_t npix = geo.npix();
  const int HP = ctx.HP;
  const int L = p->L, ncell = p->ncell;
  // loss scale for the 16-bit gradient operands (device side, no sync) and the head bias gradient (sum of dlogit,
  // unscaled): one pass over dy and y
  CU_TRY(cudaMemsetAsync(ctx.amax, 0, 16, st));  // [0] max |dlogit|, [1] max |S * dz| (float bits)
  {
    const size_t per_b = static_cast<size_t>(c.t_out) * c.height * c.width;
    dim3 grid(p->hb_chunks, c.batch * c.out_channels);
    head_grad_stats_kernel<<<grid, 256, 0, st>>>(dy, y, p->hbpart, (c.grad_scale > 0.f) ? nullptr : ctx.amax, c.batch,
                                                 c.out_channels, per_b, p->hb_chunks);
    RC_TRY(after_launch("head_grad_stats_kernel"));
  }
  choose_scale_kernel<<<1, 1, 0, st>>>(ctx.amax, ctx.scale, ctx.dtype == CLSTM_F16 ? 1024.f : 1.f, c.grad_scale);
  RC_TRY(after_launch("choose_scale_kernel"));

  for (int k = 0; k < ncell; ++k) {
    p->cells[k].bwd_started = false;
    CU_TRY(cudaMemsetAsync(p->cells[k].dc, 0, npix * HP * 4, st));
  }
  if (float* db = grads[2 * ncell + 1]) {
    const int rows = c.batch * p->hb_chunks;
    reduce_rows_kernel<<<(c.out_channels + 63) / 64, 64, 0, st>>>(p->hbpart, db, c.out_channels, rows, 1, rows, 1.f,
                                                                  accumulate);
    RC_TRY(after_launch("reduce_rows_kernel"));
  }

  // Backward schedule.  Per cell step: G = gate-grad (HBM bound) -> D = dgrad (tensor bound, on the critical
  // chain) and W = wgrad (tensor bound, off the chain: only the final reduction needs it).  W runs on a side
  // stream and is released when D has been issued, so W of step n executes concurrently with G of step n+1
  // (the gate-grad kernel is small enough to co-reside with a wgrad CTA on every SM) — two dz buffers alternate.
  const bool overlap = ctx.knobs.overlap != 0 && ctx.side != nullptr;
  int nstep = 0;
  if (overlap) {
    CU_TRY(cudaEventRecord(ctx.ev_fork, st));
    CU_TRY(cudaStreamWaitEvent(ctx.side, ctx.ev_fork, 0));  // the side stream starts after everything enqueued so far
  }
  auto back = [&](int k, int t, const float* e1, const float* e2) -> int {
    CellState& cs = p->cells[k];
    const InputRef in = plan_input(p, k, t);
    const E* gates = static_cast<const E*>(cs.gates) + static_cast<size_t>(t) * npix * 4 * HP;
    const float* c_prev = (t == 0) ? nullptr : cptr(ctx, cs, cslot(cs, t));
    const float* c_next = cptr(ctx, cs, cslot(cs, t + 1));
    const float* own = (t == cs.T - 1) ? nullptr : cs.dh_own;
    if (!overlap)
      return cell_backward_step<E>(ctx, cs, in, hslot(cs, t), gates, c_prev, c_next, own, e1, e2, st,
                                   /*need_dgrad=*/cs.with_x || t > 0);
    const int buf = nstep & 1;
    const int first = cs.bwd_started ? 0 : 1;
    cs.bwd_started = true;
    if (nstep >= 2) CU_TRY(cudaStreamWaitEvent(st, ctx.ev_w[buf], 0));  // wgrad n-2 has finished reading dz[buf]
    RC_TRY(cell_gate_grad<E>(ctx, cs, gates, c_prev, c_next, own, e1, e2, first, st, buf));
    RC_TRY(cell_dgrad<E>(ctx, cs, st, buf));
    CU_TRY(cudaEventRecord(ctx.ev_d[buf], st));
    CU_TRY(cudaStreamWaitEvent(ctx.side, ctx.ev_d[buf], 0));
    RC_TRY(cell_wgrad<E>(ctx, cs, in, hslot(cs, t), first, ctx.side, buf));
    CU_TRY(cudaEventRecord(ctx.ev_w[buf], ctx.side));
    ++nstep;
    return 0;
  };

  CellState& last = p->cells[ncell - 1];
  // head backward for output frame t: dlogit "col" tensor -> dgrad into dstack, wgrad accumulation
  auto head_back = [&](int t, bool with_dgrad = true) -> int {
    bool tiled = false;
    RC_TRY((launch_row_im2col<E, 1>(dy, y, static_cast<E*>(p->G), c.batch, c.t_out, c.out_channels, c.height, c.width,
                                    3, 3, p->KG, t, 1, ctx.scale, st, &tiled)));
    if (!tiled) {
      head_grad_col_kernel<E><<<kPackBlocks, 256, 0, st>>>(dy, y, static_cast<E*>(p->G), c.batch, c.out_channels,
                                                           c.t_out, c.height, c.width, p->KG, t, 1, ctx.scale);
      RC_TRY(after_launch("head_grad_col_kernel"));
    }
    if (with_dgrad) {
      ConvGemmParams hp;
      memset(&hp, 0, sizeof(hp));
      hp.n_tile = p->n_tile_hd;
      hp.n_tiles = HP / p->n_tile_hd;
      hp.nseg = 1;
      hp.seg[0] = ConvSeg{p->KG / 64, 1, 1, 0};
      hp.out0 = p->dstack, hp.out1 = p->dstack;
      hp.split_col = HP, hp.ld0 = HP, hp.ld1 = HP;
      hp.out_scale = 1.f;
      RC_TRY((launch_convgemm<E, EPI_STORE>(ctx, p->m_G128, p->m_G128, p->m_whd, hp, geo, c.batch, st,
                                            &p->m_dstack16, &p->m_dstack16, &p->m_dstack16, nullptr, "head_dgrad")));
    }
    {
      WgradParams wp;
      memset(&wp, 0, sizeof(wp));
      wp.n_blocks = p->KG / 128;
      wp.group_size = p->head_group;
      wp.total_blocks = HP / 64;
      wp.seg[0] = WgradSeg{HP / 64, HP / 64, 1, 0, 0, hslot(last, t + 1) * c.batch};
      wp.seg[1] = WgradSeg{0, 1, 1, 0, 0, 0};
      wp.splits = p->head_splits;
      wp.partial = p->hpart;
      wp.accumulate = (t != c.t_out - 1);
      RC_TRY((launch_wgrad<E>(ctx, p->m_G64, last.m_h64, last.m_h64, wp, geo, c.batch, st, nullptr, "head_wgrad")));
    }
    return 0;
  };

  // The chain of cell steps in backward order (conv_lstm.py:176-196 reversed); e1 / e2 = the dh sources besides the
  // cell's own recurrent dh: the head (dstack), the cell above (its dx) and the decoder feedback (dx of cell L).
  struct BackOp {
    int k, t;
    const float *e1, *e2;
    bool head;  // the head backward of frame t must have run before this step's gate gradient
  };
  std::vector<BackOp> ops;
  ops.reserve(static_cast<size_t>(L) * (c.t_in + c.t_out));
  for (int t = c.t_out - 1; t >= 0; --t) {
    ops.push_back({ncell - 1, t, p->dstack, (t == c.t_out - 1) ? nullptr : p->cells[L].dxb, true});
    for (int k = ncell - 2; k >= L; --k) ops.push_back({k, t, p->cells[k + 1].dxb, nullptr, false});
  }
  for (int t = c.t_in - 1; t >= 0; --t) {
    ops.push_back({L - 1, t, (t == c.t_in - 1) ? p->cells[L].dxb : nullptr, nullptr, false});
    for (int k = L - 2; k >= 0; --k) ops.push_back({k, t, p->cells[k + 1].dxb, nullptr, false});
  }

  // Fused schedule: the gate gradient of step n+1 runs inside the epilogue of dgrad n (dgradT.cuh) whenever the shapes
  // allow, so the HBM-bound pointwise pass overlaps the tensor-bound GEMM.  dz alternates between two buffers
  // (dgrad n reads dz[b] through TMA while its epilogue writes dz[b^1]).
  const bool fuse = !overlap && L >= 2 && fuse_supported(ctx);
  // Experimental alternative (CLSTM_WG_GATE=1, off): the gate gradient of step n+1 rides along with WGRAD n instead
  // of the dgrad epilogue — 16 extra warps in the wgrad CTAs (wgrad.cuh), plain dgradT writes dx again.
  const bool wg_gate = !overlap && HP == 64 && npix * 4 * HP < (1ull << 32) && ctx.knobs.wg_gate;
  int bias_rows = kGateGradBlocks;
  if (wg_gate) {
    bias_rows = kBiasRowsMax;
    for (int k = 0; k < ncell; ++k)
      CU_TRY(cudaMemsetAsync(p->cells[k].bpart, 0, static_cast<size_t>(kBiasRowsMax) * 4 * HP * 4, st));
    int b = 0;
    bool gate_done = false;
    for (size_t n = 0; n < ops.size(); ++n) {
      const BackOp& o = ops[n];
      CellState& cs = p->cells[o.k];
      const InputRef in = plan_input(p, o.k, o.t);
      const int first = cs.bwd_started ? 0 : 1;
      cs.bwd_started = true;
      if (!gate_done) {
        if (o.head) RC_TRY(head_back(o.t));
        const E* gates = static_cast<const E*>(cs.gates) + static_cast<size_t>(o.t) * npix * 4 * HP;
        const float* c_prev = (o.t == 0) ? nullptr : cptr(ctx, cs, cslot(cs, o.t));
        const float* c_next = cptr(ctx, cs, cslot(cs, o.t + 1));
        const float* own = (o.t == cs.T - 1) ? nullptr : cs.dh_own;
        RC_TRY(cell_gate_grad<E>(ctx, cs, gates, c_prev, c_next, own, o.e1, o.e2, 0, st, b));
      }
      gate_done = false;
      if (cs.with_x || o.t > 0) RC_TRY(cell_dgrad<E>(ctx, cs, st, b));
      if (n + 1 < ops.size()) {
        const BackOp& nx = ops[n + 1];
        CellState& cn = p->cells[nx.k];
        if (nx.head) RC_TRY(head_back(nx.t));
        WgGateWork gw;
        memset(&gw, 0, sizeof(gw));
        gw.gates = static_cast<const E*>(cn.gates) + static_cast<size_t>(nx.t) * npix * 4 * HP;
        gw.c_prev = (nx.t == 0) ? nullptr : cn.c + static_cast<size_t>(cslot(cn, nx.t)) * npix * HP;
        gw.c_next = cn.c + static_cast<size_t>(cslot(cn, nx.t + 1)) * npix * HP;
        gw.src0 = (nx.t == cn.T - 1) ? nullptr : cn.dh_own;
        gw.src1 = nx.e1, gw.src2 = nx.e2;
        gw.dc = cn.dc;
        gw.dz_out = ctx.dzb[b ^ 1];
        gw.bias_partial = cn.bpart;
        gw.dz_absmax = ctx.amax + 1;
        gw.npix = static_cast<unsigned>(npix);
        RC_TRY(cell_wgrad<E>(ctx, cs, in, hslot(cs, o.t), first, st, b, &gw));
        gate_done = true;
        b ^= 1;
      } else {
        RC_TRY(cell_wgrad<E>(ctx, cs, in, hslot(cs, o.t), first, st, b));
      }
    }
  } else if (fuse) {
    // hybrid split (see below): only when a unit is 256 CONSECUTIVE pixels (one-row 128-pixel tiles, W a multiple
    // of 256), so that "units >= hybrid_units" is the pixel range [256 * hybrid_units, npix)
    int hybrid_units = 0;
    if (ctx.knobs.hybrid_pct > 0 && ctx.knobs.hybrid_pct < 100 && geo.BW == 128 && geo.BH == 1 && geo.W % 256 == 0) {
      const long long units = static_cast<long long>(npix / 256);
      hybrid_units = static_cast<int>(units * ctx.knobs.hybrid_pct / 100);
      if (hybrid_units < 1 || hybrid_units >= units) hybrid_units = 0;
    }
    if (hybrid_units > 0) bias_rows = kBiasRowsMax;
    // 16-bit recurrent gradient states (dc, own dh_prev): every dgrad of this schedule is the worker-warp fused kernel
    // (each op but the last is fused into the next one, the last one's dgrad is not needed), the only other writer /
    // reader of dc is the stand-alone gate gradient of the very first op
    bool state16 = state16_knobs_ok<E>(ctx) && hybrid_units == 0;
    for (int k = 0; k < ncell && state16; ++k)
      if (p->cells[k].with_x ? p->cells[k].g.CIP != 64 : k != 0) state16 = false;
    for (int k = 0; k < ncell; ++k)
      CU_TRY(cudaMemsetAsync(p->cells[k].bpart, 0,
                             static_cast<size_t>(hybrid_units > 0 ? kBiasRowsMax : kGateGradBlocks) * 4 * HP * 4, st));
    int b = 0;
    bool gate_done = false;  // gate gradient of ops[n] already computed into dz[b] by the previous dgrad
    for (size_t n = 0; n < ops.size(); ++n) {
      const BackOp& o = ops[n];
      CellState& cs = p->cells[o.k];
      const InputRef in = plan_input(p, o.k, o.t);
      const int first = cs.bwd_started ? 0 : 1;
      cs.bwd_started = true;
      if (!gate_done) {
        if (o.head) RC_TRY(head_back(o.t));
        const E* gates = static_cast<const E*>(cs.gates) + static_cast<size_t>(o.t) * npix * 4 * HP;
        const float* c_prev = (o.t == 0) ? nullptr : cptr(ctx, cs, cslot(cs, o.t));
        const float* c_next = cptr(ctx, cs, cslot(cs, o.t + 1));
        const float* own = (o.t == cs.T - 1) ? nullptr : cs.dh_own;
        if (state16 && n != 0) return fail(CLSTM_EINVAL, "backward schedule: an unfused step inside the 16-bit-state chain");
        RC_TRY(cell_gate_grad<E>(ctx, cs, gates, c_prev, c_next, own, o.e1, o.e2, 0, st, b, state16));
      }
      gate_done = false;
      bool fused_here = false;
      WgGateWork gw;
      const WgGateWork* gwp = nullptr;
      if (n + 1 < ops.size() && (!cs.with_x || cs.g.CIP == 64)) {
        const BackOp& nx = ops[n + 1];
        CellState& cn = p->cells[nx.k];
        if (nx.k != o.k) {
          // The head's dgrad of frame nx.t is one more contribution to the consumer's dh: instead of a launch of its
          // own writing dstack (fp32, re-read by the gate gradient), its K = KG columns are appended to this dgrad's
          // K loop (rows of the x part, i.e. the channels of the top decoder cell's h) — dstack does not exist.
          const bool head_in_gemm = nx.head && ctx.knobs.head_fuse && cs.with_x && cs.g.CIP == 64 && hybrid_units == 0 &&
                                    fused2_ok(ctx) && nx.e1 == p->dstack && nx.e2 == cs.dxb;
          if (nx.head) RC_TRY(head_back(nx.t, !head_in_gemm));  // dstack / G must be ready; nothing in flight reads them
          const float* own_n = (nx.t == cn.T - 1) ? nullptr : cn.dh_own;
          int fuse_units = 0x7fffffff;
          if (hybrid_units > 0 && cs.with_x) {
            // hybrid: the first hybrid_units units (256 consecutive pixels each) finish their gate gradient in this
            // dgrad's epilogue, the remaining pixels on the worker warps of the wgrad launch below (they read dx,
            // which the dgrad then writes for those units only).  Spreads the HBM-bound pointwise pass over BOTH GEMMs.
            fuse_units = hybrid_units;
            memset(&gw, 0, sizeof(gw));
            gw.gates = static_cast<const E*>(cn.gates) + static_cast<size_t>(nx.t) * npix * 4 * HP;
            gw.c_prev = (nx.t == 0) ? nullptr : cn.c + static_cast<size_t>(cslot(cn, nx.t)) * npix * HP;
            gw.c_next = cn.c + static_cast<size_t>(cslot(cn, nx.t + 1)) * npix * HP;
            gw.src0 = own_n, gw.src1 = nx.e1, gw.src2 = nx.e2;
            gw.dc = cn.dc;
            gw.dz_out = ctx.dzb[b ^ 1];
            gw.bias_partial = cn.bpart;
            gw.dz_absmax = ctx.amax + 1;
            gw.npix = static_cast<unsigned>(npix);
            gw.pix_begin = static_cast<unsigned>(hybrid_units) * 256u;
            gwp = &gw;
          }
          if (head_in_gemm)
            RC_TRY(cell_dgrad_fused<E>(ctx, cs, cn, nx.t, own_n, nullptr, nx.e2, b, st, fuse_units, &p->m_G128,
                                       &p->m_whdT, p->KG / 64, state16));
          else
            RC_TRY(cell_dgrad_fused<E>(ctx, cs, cn, nx.t, own_n, nx.e1, nx.e2, b, st, fuse_units, nullptr, nullptr, 0,
                                       state16));
          fused_here = true;
          gate_done = true;
        }
      }
      // the bottom cell's step 0 has nobody to hand a gradient to: its dgrad would only produce d(initial h) = unused
      if (!fused_here && (cs.with_x || o.t > 0)) {
        if (state16) return fail(CLSTM_EINVAL, "backward schedule: a plain dgrad inside the 16-bit-state chain");
        RC_TRY(cell_dgrad<E>(ctx, cs, st, b));
      }
      RC_TRY(cell_wgrad<E>(ctx, cs, in, hslot(cs, o.t), first, st, b, gwp));
      if (fused_here) b ^= 1;
    }
  } else {
    for (const BackOp& o : ops) {
      if (o.head) RC_TRY(head_back(o.t));
      RC_TRY(back(o.k, o.t, o.e1, o.e2));
    }
  }
  if (overlap) {  // join: the partial sums of every wgrad are complete
    CU_TRY(cudaEventRecord(ctx.ev_fork, ctx.side));
    CU_TRY(cudaStreamWaitEvent(st, ctx.ev_fork, 0));
  }
  for (int k = 0; k < ncell; ++k)
    RC_TRY(cell_finalize(ctx, p->cells[k], grads[2 * k], grads[2 * k + 1], accumulate, st, bias_rows));
  if (grads[2 * ncell]) {
    const int total = c.out_channels * c.hidden * 9;
    head_wgrad_finalize_kernel<<<(total + 255) / 256, 256, 0, st>>>(p->hpart, grads[2 * ncell], c.out_channels,
                                                                    c.hidden, HP, p->KG, p->head_splits,
                                                                    ctx.scale + 1, accumulate);
    RC_TRY(after_launch("head_wgrad_finalize_kernel"));
  }
  return 0;
}

template <typename E>
int plan_read_state(clstm_plan* p, int cell, int step, float* h_out, float* c_out, cudaStream_t st) {
  const clstm_config_t& c = p->cfg;
  CellState& cs = p->cells[cell];
  const size_t npix = p->ctx.geo.npix();
  const int HP = p->ctx.HP;
  if (h_out) {
    const E* src = static_cast<const E*>(cs.h) + static_cast<size_t>(hslot(cs, step)) * npix * HP;
    unpack_nchw_kernel<E><<<kPackBlocks, 256, 0, st>>>(src, h_out, c.batch, c.hidden, c.height, c.width, HP, nullptr, 0,
                                                       kHScaleInv);
    RC_TRY(after_launch("unpack_nchw_kernel"));
  }
  if (c_out) {
    if (step == 0) {
      CU_TRY(cudaMemsetAsync(c_out, 0, static_cast<size_t>(c.batch) * c.hidden * c.height * c.width * 4, st));
    } else {
      const float* src = cptr(p->ctx, cs, cslot(cs, step));
      if (p->ctx.c16)
        unpack_nchw_kernel<E><<<kPackBlocks, 256, 0, st>>>(reinterpret_cast<const E*>(src), c_out, c.batch, c.hidden,
                                                           c.height, c.width, HP, nullptr, 0, kCScaleInv);
      else
        unpack_nchw_kernel<float><<<kPackBlocks, 256, 0, st>>>(src, c_out, c.batch, c.hidden, c.height, c.width, HP,
                                                               nullptr, 0);
      RC_TRY(after_launch("unpack_nchw_kernel"));
    }
  }
  return 0;
}

}  // namespace

// ============================================================================ single-cell plan
struct clstm_cell_plan {
  Ctx ctx;
  CellState cs;
  int cin = 0, hid = 0;
  size_t saved_bytes = 0, scratch_bytes = 0;
  bool bound = false, forward_done = false;
  int cur = 0;          // native stepping: slot (0 / 1) holding the current h / c state
  bool native_ready = false;
  void* xin = nullptr;  // E [npix][CIP]
  CUtensorMap m_x128, m_x64;
};

namespace {

// saved region: packed x, packed weights, h / c (both steps), gates — everything clstm_cell_backward reads that the
// forward wrote; scratch region: loss scale, dz, dh / dx / dc, split partials.
void carve_cell_plan(clstm_cell_plan* p, uint8_t* saved, uint8_t* scratch) {
  Carver sv, sc;
  sv.base = saved, sc.base = scratch;
  Ctx& ctx = p->ctx;
  const size_t npix = ctx.geo.npix();
  ctx.scale = sc.take<float>(64);
  ctx.amax = reinterpret_cast<unsigned int*>(sc.take<float>(64));
  p->xin = sv.take<void>(npix * p->cs.g.CIP * 2);
  carve_cell(sv, sc, p->cs, ctx);
  ctx.dz = sc.take<void>(npix * 4 * ctx.HP * 2);
  ctx.dzb[0] = ctx.dzb[1] = ctx.dz;
  p->saved_bytes = align_up(sv.off, 1024);
  p->scratch_bytes = align_up(sc.off, 1024);
}

template <typename E>
int cellplan_forward(clstm_cell_plan* p, const float* x, const float* h_cur, const float* c_cur,
                     const float* weight, const float* bias, float* h_next, float* c_next, cudaStream_t st) {
  Ctx& ctx = p->ctx;
  CellState& cs = p->cs;
  const Geo& geo = ctx.geo;
  const size_t npix = geo.npix();
  const int HP = ctx.HP;
  const size_t plane = static_cast<size_t>(geo.H) * geo.W;
  RC_TRY(pack_cell<E>(ctx, cs, weight, bias, st));
  pack_nhwc_kernel<E><<<kPackBlocks, 256, 0, st>>>(x, static_cast<E*>(p->xin), geo.B, p->cin, geo.H, geo.W, cs.g.CIP,
                                                   p->cin * plane, 1.f);
  RC_TRY(after_launch("pack_nhwc_kernel"));
  pack_nhwc_kernel<E><<<kPackBlocks, 256, 0, st>>>(h_cur, static_cast<E*>(cs.h), geo.B, p->hid, geo.H, geo.W, HP,
                                                   p->hid * plane, kHScale);
  RC_TRY(after_launch("pack_nhwc_kernel"));
  pack_nhwc_f32_kernel<<<kPackBlocks, 256, 0, st>>>(c_cur, cs.c, geo.B, p->hid, geo.H, geo.W, HP, nullptr);
  RC_TRY(after_launch("pack_nhwc_f32_kernel"));
  InputRef in;
  in.map128 = &p->m_x128, in.map64 = &p->m_x64, in.b_off = 0;
  RC_TRY(cell_forward_step<E>(ctx, cs, in, 0, 1, cs.c, cs.c + npix * HP, cs.gates, st, 0, 1, 0));
  if (h_next) {
    unpack_nchw_kernel<E><<<kPackBlocks, 256, 0, st>>>(static_cast<const E*>(cs.h) + npix * HP, h_next, geo.B, p->hid,
                                                       geo.H, geo.W, HP, nullptr, 0, kHScaleInv);
    RC_TRY(after_launch("unpack_nchw_kernel"));
  }
  if (c_next) {
    unpack_nchw_kernel<float><<<kPackBlocks, 256, 0, st>>>(cs.c + npix * HP, c_next, geo.B, p->hid, geo.H, geo.W, HP,
                                                           nullptr, 0);
    RC_TRY(after_launch("unpack_nchw_kernel"));
  }
  p->forward_done = true;
  return 0;
}

// ---- native-layout stepping (clstm_cell_native_*): the state stays in the plan's NHWC 16-bit / fp32 slots between
// steps, the weights are packed once, so one step is exactly one fused kernel (no NCHW <-> NHWC packing).
template <typename E>
int cellplan_native_load(clstm_cell_plan* p, const float* x, const float* h, const float* c, const float* weight,
                         const float* bias, int reset_state, cudaStream_t st) {
  Ctx& ctx = p->ctx;
  CellState& cs = p->cs;
  const Geo& geo = ctx.geo;
  const size_t npix = geo.npix();
  const int HP = ctx.HP;
  const size_t plane = static_cast<size_t>(geo.H) * geo.W;
  if (weight) RC_TRY(pack_cell<E>(ctx, cs, weight, bias, st));
  if (x) {
    pack_nhwc_kernel<E><<<kPackBlocks, 256, 0, st>>>(x, static_cast<E*>(p->xin), geo.B, p->cin, geo.H, geo.W, cs.g.CIP,
                                                     p->cin * plane, 1.f);
    RC_TRY(after_launch("pack_nhwc_kernel"));
  }
  E* hcur = static_cast<E*>(cs.h) + static_cast<size_t>(p->cur) * npix * HP;
  float* ccur = cs.c + static_cast<size_t>(p->cur) * npix * HP;
  if (h) {
    pack_nhwc_kernel<E><<<kPackBlocks, 256, 0, st>>>(h, hcur, geo.B, p->hid, geo.H, geo.W, HP, p->hid * plane, kHScale);
    RC_TRY(after_launch("pack_nhwc_kernel"));
  } else if (reset_state) {
    CU_TRY(cudaMemsetAsync(hcur, 0, npix * HP * sizeof(E), st));  // ConvLSTMCell.init_hidden (layers/ConvLSTM.py:59-64)
  }
  if (c) {
    pack_nhwc_f32_kernel<<<kPackBlocks, 256, 0, st>>>(c, ccur, geo.B, p->hid, geo.H, geo.W, HP, nullptr);
    RC_TRY(after_launch("pack_nhwc_f32_kernel"));
  } else if (reset_state) {
    CU_TRY(cudaMemsetAsync(ccur, 0, npix * HP * 4, st));
  }
  return 0;
}

template <typename E>
int cellplan_native_step(clstm_cell_plan* p, cudaStream_t st) {
  Ctx& ctx = p->ctx;
  CellState& cs = p->cs;
  const size_t npix = ctx.geo.npix();
  const int HP = ctx.HP;
  InputRef in;
  in.map128 = &p->m_x128, in.map64 = &p->m_x64, in.b_off = 0;
  const int a = p->cur, b = p->cur ^ 1;
  // gates == nullptr: nothing is kept for a backward, the step writes only h' and c'
  RC_TRY(cell_forward_step<E>(ctx, cs, in, a, b, cs.c + static_cast<size_t>(a) * npix * HP,
                              cs.c + static_cast<size_t>(b) * npix * HP, nullptr, st, a, b, -1));
  p->cur = b;
  return 0;
}

template <typename E>
int cellplan_native_read(clstm_cell_plan* p, float* h_out, float* c_out, cudaStream_t st) {
  Ctx& ctx = p->ctx;
  CellState& cs = p->cs;
  const Geo& geo = ctx.geo;
  const size_t npix = geo.npix();
  const int HP = ctx.HP;
  if (h_out) {
    unpack_nchw_kernel<E><<<kPackBlocks, 256, 0, st>>>(static_cast<const E*>(cs.h) + static_cast<size_t>(p->cur) * npix * HP,
                                                       h_out, geo.B, p->hid, geo.H, geo.W, HP, nullptr, 0, kHScaleInv);
    RC_TRY(after_launch("unpack_nchw_kernel"));
  }
  if (c_out) {
    unpack_nchw_kernel<float><<<kPackBlocks, 256, 0, st>>>(cs.c + static_cast<size_t>(p->cur) * npix * HP, c_out, geo.B,
                                                           p->hid, geo.H, geo.W, HP, nullptr, 0);
    RC_TRY(after_launch("unpack_nchw_kernel"));
  }
  return 0;
}

template <typename E>
int cellplan_backward(clstm_cell_plan* p, const float* dh_next, const float* dc_next, const float* weight,
                      float* dx, float* dh_cur, float* dc_cur, float* dweight, float* dbias, cudaStream_t st) {
  Ctx& ctx = p->ctx;
  CellState& cs = p->cs;
  const Geo& geo = ctx.geo;
  const size_t npix = geo.npix();
  const int HP = ctx.HP;
  (void)weight;  // the packed copies made by the preceding forward are used
  // Loss scale for the 16-bit dz operand: S = 2^k with S * max(|dh_next|, |dc_next|) ~ 2^10 (device side).
  const size_t nstate = static_cast<size_t>(geo.B) * p->hid * geo.H * geo.W;
  CU_TRY(cudaMemsetAsync(ctx.amax, 0, 16, st));
  if (dh_next) {
    abs_amax_kernel<<<kPackBlocks, 256, 0, st>>>(dh_next, nstate, ctx.amax);
    RC_TRY(after_launch("abs_amax_kernel"));
  }
  if (dc_next) {
    abs_amax_kernel<<<kPackBlocks, 256, 0, st>>>(dc_next, nstate, ctx.amax);
    RC_TRY(after_launch("abs_amax_kernel"));
  }
  choose_scale_kernel<<<1, 1, 0, st>>>(ctx.amax, ctx.scale, ctx.dtype == CLSTM_F16 ? 1024.f : 1.f, 0.f);
  RC_TRY(after_launch("choose_scale_kernel"));
  // upstream gradients -> NHWC fp32, scaled (dh into dh_own as source 0, dc into the in-place dc buffer)
  float* dh_src = nullptr;
  if (dh_next) {
    pack_nhwc_f32_kernel<<<kPackBlocks, 256, 0, st>>>(dh_next, cs.dh_own, geo.B, p->hid, geo.H, geo.W, HP, ctx.scale);
    RC_TRY(after_launch("pack_nhwc_f32_kernel"));
    dh_src = cs.dh_own;
  }
  if (dc_next) {
    pack_nhwc_f32_kernel<<<kPackBlocks, 256, 0, st>>>(dc_next, cs.dc, geo.B, p->hid, geo.H, geo.W, HP, ctx.scale);
    RC_TRY(after_launch("pack_nhwc_f32_kernel"));
  } else {
    CU_TRY(cudaMemsetAsync(cs.dc, 0, npix * HP * 4, st));
  }
  cs.bwd_started = false;
  InputRef in;
  in.map128 = &p->m_x128, in.map64 = &p->m_x64, in.b_off = 0;
  RC_TRY(cell_backward_step<E>(ctx, cs, in, 0, cs.gates, cs.c, cs.c + npix * HP, dh_src, nullptr, nullptr, st));
  RC_TRY(cell_finalize(ctx, cs, dweight, dbias, 0, st));
  if (dx) {
    unpack_nchw_kernel<float><<<kPackBlocks, 256, 0, st>>>(cs.dxb, dx, geo.B, p->cin, geo.H, geo.W, cs.g.CIP,
                                                           ctx.scale + 1, 0);
    RC_TRY(after_launch("unpack_nchw_kernel"));
  }
  if (dh_cur) {
    unpack_nchw_kernel<float><<<kPackBlocks, 256, 0, st>>>(cs.dh_own, dh_cur, geo.B, p->hid, geo.H, geo.W, HP,
                                                           ctx.scale + 1, 0);
    RC_TRY(after_launch("unpack_nchw_kernel"));
  }
  if (dc_cur) {
    unpack_nchw_kernel<float><<<kPackBlocks, 256, 0, st>>>(cs.dc, dc_cur, geo.B, p->hid, geo.H, geo.W, HP,
                                                           ctx.scale + 1, 0);
    RC_TRY(after_launch("unpack_nchw_kernel"));
  }
  return 0;
}

}  // namespace

// ============================================================================ C ABI
extern "C" {

const char* clstm_last_error(void) { return g_err; }
int clstm_abi_version(void) { return CLSTM_ABI_VERSION; }
uint64_t clstm_launch_count(void) { return g_launches.load(); }

int clstm_trace_enable(int capacity) {
  for (cudaEvent_t e : g_trace.ev) cudaEventDestroy(e);
  g_trace.ev.clear();
  g_trace.name.clear();
  g_trace.n = 0;
  g_trace.on = false;
  if (capacity <= 0) return 0;
  g_trace.ev.resize(static_cast<size_t>(capacity));
  g_trace.name.assign(static_cast<size_t>(capacity), nullptr);
  for (auto& e : g_trace.ev) CU_TRY(cudaEventCreate(&e));
  g_trace.on = true;
  return 0;
}

long long clstm_trace_report(char* buf, size_t cap) {
  if (!buf || cap == 0) return fail(CLSTM_EINVAL, "null argument");
  buf[0] = 0;
  const size_t n = g_trace.n;
  if (n < 2) return 0;
  CU_TRY(cudaEventSynchronize(g_trace.ev[n - 1]));
  struct Row {
    const char* name;
    long long count;
    double ms;
  };
  std::vector<Row> rows;
  double total = 0.0;
  for (size_t i = 1; i < n; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_trace.ev[i - 1], g_trace.ev[i]) != cudaSuccess) continue;
    const char* nm = g_trace.name[i] ? g_trace.name[i] : "(outside the library: between API calls)";
    Row* r = nullptr;
    for (auto& x : rows)
      if (strcmp(x.name, nm) == 0) r = &x;
    if (!r) {
      rows.push_back(Row{nm, 0, 0.0});
      r = &rows.back();
    }
    r->count += 1;
    r->ms += ms;
    total += ms;
  }
  for (size_t i = 0; i < rows.size(); ++i)  // sort by total time, descending
    for (size_t j = i + 1; j < rows.size(); ++j)
      if (rows[j].ms > rows[i].ms) std::swap(rows[i], rows[j]);
  size_t off = 0;
  auto put = [&](const char* fmt, ...) {
    if (off >= cap - 1) return;
    va_list ap;
    va_start(ap, fmt);
    const int w = vsnprintf(buf + off, cap - off, fmt, ap);
    va_end(ap);
    if (w > 0) off += static_cast<size_t>(w) < cap - off ? static_cast<size_t>(w) : cap - off - 1;
  };
  put("%zu events, %.3f ms between the first and the last%s\n", n, total,
      n >= g_trace.ev.size() ? " (capacity reached: later launches were dropped)" : "");
  for (const Row& r : rows)
    put("%10.3f ms %5.1f%%  n=%5lld  avg=%9.1f us  %s\n", r.ms, 100.0 * r.ms / (total > 0 ? total : 1), r.count,
        1e3 * r.ms / r.count, r.name);
  g_trace.n = 0;
  return static_cast<long long>(off);
}

int clstm_device_check(int ordinal) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) return fail(CLSTM_ENODEV, "no CUDA device: %s", cudaGetErrorString(e));
  if (ordinal < 0 || ordinal >= n) return fail(CLSTM_ENODEV, "device ordinal %d out of range (%d devices)", ordinal, n);
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, ordinal));
  if (prop.major != 10)
    return fail(CLSTM_ENODEV, "device %d (%s) is sm_%d%d; libclstm is sm_100a only", ordinal, prop.name, prop.major,
                prop.minor);
  return 0;
}

int clstm_plan_create(const clstm_config_t* cfg, clstm_plan_t** out) {
  if (!cfg || !out) return fail(CLSTM_EINVAL, "null argument");
  *out = nullptr;
  RC_TRY(validate_common(cfg->batch, cfg->height, cfg->width, cfg->in_channels, cfg->hidden, cfg->kernel_h,
                         cfg->kernel_w, cfg->dtype));
  if (cfg->out_channels < 1 || cfg->out_channels > 256) return fail(CLSTM_EINVAL, "out_channels must be in [1,256]");
  if (cfg->n_layers < 1 || cfg->n_layers > 8) return fail(CLSTM_EINVAL, "n_layers must be in [1,8]");
  if (cfg->t_in < 1) return fail(CLSTM_EINVAL, "t_in must be >= 1");
  if (cfg->t_out < 1)
    return fail(CLSTM_EINVAL,
                "t_out (forecast_steps) must be >= 1: the reference raises on an empty output stack "
                "(conv_lstm.py:198)");
  clstm_plan* p = new (std::nothrow) clstm_plan();
  if (!p) return fail(CLSTM_EINVAL, "out of host memory");
  p->cfg = *cfg;
  Ctx& ctx = p->ctx;
  // Device properties are optional at create time (sizes do not depend on them except the wgrad split,
  // which falls back to 148 SMs); bind re-checks the device.
  DeviceInfo dev;
  if (get_device(&dev) == 0) ctx.dev = dev;
  ctx.geo = make_geo(cfg->batch, cfg->height, cfg->width);
  ctx.knobs.read();
  ctx.dtype = cfg->dtype;
  ctx.HP = pad_hidden(cfg->hidden);
  ctx.training = cfg->training;
  ctx.grad_scale = cfg->grad_scale;
  p->L = cfg->n_layers;
  p->ncell = 2 * cfg->n_layers;
  {
    // 16-bit c stacks (ptx.cuh kCScale): big fp16 rollouts whose every cell step takes the staged TMA epilogue (not the
    // persistent chain, which keeps c in TMEM and streams it out as fp32) and whose backward reads c through kernels
    // that know the format: the stand-alone gate gradient, or the worker-warp fused dgrad with c' recomputed.
    const Knobs& k = ctx.knobs;
    const long long tiles = static_cast<long long>(ctx.geo.B) * ctx.geo.tiles_w * ctx.geo.tiles_h * (ctx.HP / 64);
    const int sms = ctx.dev.sms > 0 ? ctx.dev.sms : 148;
    const bool fused_chain = k.fuse_gate && k.dgradT && ctx.HP == 64 && cfg->n_layers >= 2 && !k.overlap;
    ctx.c16 = k.c16 && cfg->dtype == CLSTM_F16 && k.staged == 1 && cfg->t_in <= kC16MaxSteps &&
              cfg->t_out <= kC16MaxSteps && tiles > sms && !k.wg_gate && k.hybrid_pct == 0 &&
              (!cfg->training || !fused_chain || (k.recomp_c && k.fuse_workers == 2));
  }
  p->KX = round_up(cfg->kernel_h * cfg->kernel_w * cfg->in_channels, 64);
  p->KG = round_up(9 * cfg->out_channels, 128);
  p->NT = round_up(cfg->out_channels, 16);
  p->cells.resize(p->ncell);
  for (int k = 0; k < p->ncell; ++k) {
    CellState& cs = p->cells[k];
    const int T = (k < p->L) ? cfg->t_in : cfg->t_out;
    init_cell(&cs, ctx, k == 0 ? cfg->in_channels : cfg->hidden, cfg->hidden, cfg->kernel_h, cfg->kernel_w,
              k == 0 ? 1 : 0, k == 0 ? 0 : 1, T, k == 0 ? 0 : 1);  // every cell but the first reads another cell's h
    const bool full = cfg->training || k == p->ncell - 1;  // the head reads every last-decoder h
    cs.slots_h = full ? T + 1 : 2;
    cs.slots_c = cfg->training ? T + 1 : 2;
  }
  int nt = 256;
  while (ctx.HP % nt) nt -= 64;
  p->n_tile_hd = nt;
  const long long p_tiles = static_cast<long long>(ctx.geo.B) * ctx.geo.tiles_w2 * ctx.geo.tiles_h2;
  wgrad_shape(ctx.dev, ctx.knobs.wg_group, ctx.HP / 64, p->KG / 128, p_tiles, &p->head_group, &p->head_splits);
  carve_plan(p, nullptr);
  *out = p;
  return 0;
}

int clstm_plan_destroy(clstm_plan_t* plan) {
  if (plan) {
    plan->g_fwd.clear();
    plan->g_bwd.clear();
  }
  if (plan && plan->ctx.side) {
    cudaStreamSynchronize(plan->ctx.side);
    for (int i = 0; i < 2; ++i) {
      cudaEventDestroy(plan->ctx.ev_d[i]);
      cudaEventDestroy(plan->ctx.ev_w[i]);
    }
    cudaEventDestroy(plan->ctx.ev_fork);
    cudaStreamDestroy(plan->ctx.side);
  }
  delete plan;
  return 0;
}

size_t clstm_plan_workspace_bytes(const clstm_plan_t* plan) { return plan ? plan->ws_bytes : 0; }

int clstm_plan_bind(clstm_plan_t* p, void* workspace, size_t bytes, void* stream) {
  if (!p || !workspace) return fail(CLSTM_EINVAL, "null argument");
  if (bytes < p->ws_bytes) return fail(CLSTM_EINVAL, "workspace too small: %zu < %zu", bytes, p->ws_bytes);
  if (reinterpret_cast<uintptr_t>(workspace) % 1024) return fail(CLSTM_EINVAL, "workspace must be 1024-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Ctx& ctx = p->ctx;
  DeviceInfo dev;
  RC_TRY(get_device(&dev));
  if (ctx.dev.sms != 0 && ctx.dev.sms != dev.sms)
    return fail(CLSTM_ESTATE, "plan was created for a device with %d SMs, bound on one with %d", ctx.dev.sms, dev.sms);
  if (ctx.dev.sms == 0 && dev.sms != 148)
    return fail(CLSTM_ESTATE, "plan was created without a device and assumed 148 SMs; this device has %d", dev.sms);
  ctx.dev = dev;
  carve_plan(p, static_cast<uint8_t*>(workspace));
  const Geo& g = ctx.geo;
  const clstm_config_t& c = p->cfg;
  const size_t npix = g.npix();
  // zero the initial states (ConvLSTMCell.init_hidden, layers/ConvLSTM.py:59-64) and everything a
  // tensor map may read before it is written
  for (int k = 0; k < p->ncell; ++k) {
    CellState& cs = p->cells[k];
    CU_TRY(cudaMemsetAsync(static_cast<uint8_t*>(cs.h) + static_cast<size_t>(hslot(cs, 0)) * npix * ctx.HP * 2, 0,
                           npix * ctx.HP * 2, st));
    CU_TRY(cudaMemsetAsync(cs.c, 0, npix * ctx.HP * (ctx.c16 ? 2 : 4), st));
    RC_TRY(map_cell(cs, ctx));
  }
  const long long ximgs = static_cast<long long>(c.t_in) * c.batch;
  RC_TRY(make_map_act(&p->m_xcol128, ctx.dtype, p->xcol, p->KX, g.W, g.H, ximgs, g.BW, g.BH));
  RC_TRY(make_map_act(&p->m_xcol64, ctx.dtype, p->xcol, p->KX, g.W, g.H, ximgs, g.BW2, g.BH2));
  RC_TRY(make_map_w(&p->m_wh, ctx.dtype, p->wh, 9 * ctx.HP, p->NT, p->NT));
  if (p->wz) RC_TRY(make_map_w(&p->m_wz, ctx.dtype, p->wz, 64, (ctx.HP / 64) * kHrN, kHrN));
  if (c.training) {
    RC_TRY(make_map_act(&ctx.m_dz128, ctx.dtype, ctx.dz, 4 * ctx.HP, g.W, g.H, c.batch, g.BW, g.BH));
    RC_TRY(make_map_act(&ctx.m_dz64, ctx.dtype, ctx.dz, 4 * ctx.HP, g.W, g.H, c.batch, g.BW2, g.BH2));
    for (int i = 0; i < 2; ++i) {
      RC_TRY(make_map_act(&ctx.m_dz128b[i], ctx.dtype, ctx.dzb[i], 4 * ctx.HP, g.W, g.H, c.batch, g.BW, g.BH));
      RC_TRY(make_map_act(&ctx.m_dz64b[i], ctx.dtype, ctx.dzb[i], 4 * ctx.HP, g.W, g.H, c.batch, g.BW2, g.BH2));
    }
    if (!ctx.side) {
      CU_TRY(cudaStreamCreateWithFlags(&ctx.side, cudaStreamNonBlocking));
      for (int i = 0; i < 2; ++i) {
        CU_TRY(cudaEventCreateWithFlags(&ctx.ev_d[i], cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&ctx.ev_w[i], cudaEventDisableTiming));
      }
      CU_TRY(cudaEventCreateWithFlags(&ctx.ev_fork, cudaEventDisableTiming));
    }
    RC_TRY(make_map_act(&p->m_G128, ctx.dtype, p->G, p->KG, g.W, g.H, c.batch, g.BW, g.BH));
    RC_TRY(make_map_act(&p->m_G64, ctx.dtype, p->G, p->KG, g.W, g.H, c.batch, g.BW2, g.BH2));
    RC_TRY(make_map_w(&p->m_whd, ctx.dtype, p->whd, p->KG, ctx.HP, p->n_tile_hd));
    RC_TRY(make_map_w(&p->m_whdT, ctx.dtype, p->whdT, p->KG, 128, 128));
    RC_TRY(make_map_epi(&p->m_dstack16, 4, ctx.dtype, p->dstack, ctx.HP, g.W, g.H, c.batch, g.BW, g.BH));
  }
  RC_TRY(persist_setup(p, st));
  // launch-bound when a cell step is at most a couple of waves of tiles; the side-stream schedule is not captured
  p->g_fwd.clear();
  p->g_bwd.clear();
  p->graph_ok = ctx.knobs.graph && !ctx.knobs.overlap &&
                static_cast<long long>(g.B) * g.tiles_w * g.tiles_h * (ctx.HP / 64) <= 4ll * ctx.dev.sms;
  p->bound = true;
  p->weights_set = false;
  p->forward_done = false;
  return 0;
}

int clstm_plan_set_weights(clstm_plan_t* p, const float* const* params, int n_params, void* stream) {
  if (!p || !params) return fail(CLSTM_EINVAL, "null argument");
  if (!p->bound) return fail(CLSTM_ESTATE, "clstm_plan_set_weights before clstm_plan_bind");
  if (n_params != 2 * p->ncell + 2)
    return fail(CLSTM_EINVAL, "expected %d parameter tensors, got %d", 2 * p->ncell + 2, n_params);
  for (int i = 0; i < n_params; ++i)
    if (!params[i]) return fail(CLSTM_EINVAL, "parameter %d is null", i);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CALL_(E) plan_set_weights<E>(p, params, st)
  RC_TRY(DISPATCH_E(p->cfg.dtype, CALL_));
#undef CALL_
  p->weights_set = true;
  return 0;
}

int clstm_rollout_forward(clstm_plan_t* p, const float* x, float* y, void* stream) {
  return clstm_rollout_forward_layout(p, x, CLSTM_X_BTCHW, y, stream);
}

int clstm_rollout_forward_layout(clstm_plan_t* p, const float* x, int x_layout, float* y, void* stream) {
  trace_begin(stream);
  if (!p || !x || !y) return fail(CLSTM_EINVAL, "null argument");
  if (!p->bound || !p->weights_set) return fail(CLSTM_ESTATE, "forward before bind / set_weights");
  if (x_layout != CLSTM_X_BTCHW && x_layout != CLSTM_X_BTHWC) return fail(CLSTM_EINVAL, "unknown x layout %d", x_layout);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return run_graphed(p->g_fwd, p->graph_ok, {x, y, reinterpret_cast<const void*>(static_cast<uintptr_t>(x_layout))}, st,
                     [&](cudaStream_t s) -> int {
#define CALL_(E) plan_forward<E>(p, x, y, s, x_layout == CLSTM_X_BTHWC)
                       return DISPATCH_E(p->cfg.dtype, CALL_);
#undef CALL_
                     });
}

int clstm_rollout_backward(clstm_plan_t* p, const float* dy, const float* y, float* const* grads, int n_grads,
                           int accumulate, void* stream) {
  trace_begin(stream);
  if (!p || !dy || !y || !grads) return fail(CLSTM_EINVAL, "null argument");
  if (!p->cfg.training) return fail(CLSTM_ESTATE, "backward on a plan created with training = 0");
  if (!p->forward_done) return fail(CLSTM_ESTATE, "backward before forward");
  if (n_grads != 2 * p->ncell + 2)
    return fail(CLSTM_EINVAL, "expected %d gradient tensors, got %d", 2 * p->ncell + 2, n_grads);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  std::vector<const void*> key = {dy, y, reinterpret_cast<const void*>(static_cast<uintptr_t>(accumulate != 0))};
  for (int i = 0; i < n_grads; ++i) key.push_back(grads[i]);
  return run_graphed(p->g_bwd, p->graph_ok, std::move(key), st, [&](cudaStream_t s) -> int {
#define CALL_(E) plan_backward<E>(p, dy, y, grads, accumulate, s)
    return DISPATCH_E(p->cfg.dtype, CALL_);
#undef CALL_
  });
}

int clstm_plan_grad_status(clstm_plan_t* p, float* out4, void* stream) {
  if (!p || !out4) return fail(CLSTM_EINVAL, "null argument");
  if (!p->bound || !p->cfg.training) return fail(CLSTM_ESTATE, "grad_status needs a bound training plan");
  grad_status_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(p->ctx.scale, p->ctx.amax, out4);
  return after_launch("grad_status_kernel");
}

int clstm_plan_info(const clstm_plan_t* p, int what, long long* out) {
  if (!p || !out) return fail(CLSTM_EINVAL, "null argument");
  switch (what) {
    case CLSTM_INFO_PERSISTENT_CHAIN: *out = p->persist_ok ? 1 : 0; return 0;
    case CLSTM_INFO_GRAPH_ENABLED: *out = (p->graph_ok && !p->g_fwd.disabled && !p->g_bwd.disabled) ? 1 : 0; return 0;
    case CLSTM_INFO_GRAPH_CAPTURES: *out = static_cast<long long>(p->g_fwd.captures + p->g_bwd.captures); return 0;
    case CLSTM_INFO_GRAPH_REPLAYS: *out = static_cast<long long>(p->g_fwd.replays + p->g_bwd.replays); return 0;
    default: return fail(CLSTM_EINVAL, "unknown info item %d", what);
  }
}

int clstm_plan_read_state(clstm_plan_t* p, int cell, int step, float* h_out, float* c_out, void* stream) {
  if (!p) return fail(CLSTM_EINVAL, "null argument");
  if (!p->forward_done) return fail(CLSTM_ESTATE, "read_state before forward");
  if (cell < 0 || cell >= p->ncell) return fail(CLSTM_EINVAL, "cell index %d out of range", cell);
  const CellState& cs = p->cells[cell];
  if (step < 0 || step > cs.T) return fail(CLSTM_EINVAL, "step %d out of range [0,%d]", step, cs.T);
  if (h_out && cs.slots_h < cs.T + 1 && step < cs.T - 1)
    return fail(CLSTM_ESTATE, "inference plans keep only the last two h steps of cell %d", cell);
  if (c_out && cs.slots_c < cs.T + 1 && step < cs.T - 1 && step != 0)
    return fail(CLSTM_ESTATE, "inference plans keep only the last two c steps of cell %d", cell);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CALL_(E) plan_read_state<E>(p, cell, step, h_out, c_out, st)
  return DISPATCH_E(p->cfg.dtype, CALL_);
#undef CALL_
}

int clstm_plan_profile_kernel(clstm_plan_t* p, int kind, int cell, int step, void* stream) {
  if (!p) return fail(CLSTM_EINVAL, "null argument");
  if (!p->forward_done) return fail(CLSTM_ESTATE, "profile_kernel before forward");
  if (cell < 0 || cell >= p->ncell) return fail(CLSTM_EINVAL, "cell index %d out of range", cell);
  CellState& cs = p->cells[cell];
  if (step < 0 || step >= cs.T) return fail(CLSTM_EINVAL, "step %d out of range [0,%d)", step, cs.T);
  if (kind != CLSTM_KERNEL_CELL_FWD && !p->cfg.training)
    return fail(CLSTM_ESTATE, "backward kernels need a training plan");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t npix = p->ctx.geo.npix();
  const int HP = p->ctx.HP;
  const InputRef in = plan_input(p, cell, step);
  const float* c_prev = (step == 0) ? nullptr : cptr(p->ctx, cs, cslot(cs, step));
  float* c_next = cptr(p->ctx, cs, cslot(cs, step + 1));
  void* gates = nullptr;
  if (p->cfg.training) gates = static_cast<uint8_t*>(cs.gates) + static_cast<size_t>(step) * npix * 4 * HP * 2;
  switch (kind) {
    case CLSTM_KERNEL_CELL_FWD: {
#define CALL_(E)                                                                                              \
  cell_forward_step<E>(p->ctx, cs, in, hslot(cs, step), hslot(cs, step + 1), c_prev, c_next, gates, st,               \
                       cslot(cs, step), cslot(cs, step + 1), step)
      return DISPATCH_E(p->cfg.dtype, CALL_);
#undef CALL_
    }
    case CLSTM_KERNEL_GATE_GRAD: {
#define CALL_(E) cell_gate_grad<E>(p->ctx, cs, gates, c_prev, c_next, cs.dh_own, p->dstack, nullptr, 0, st)
      return DISPATCH_E(p->cfg.dtype, CALL_);
#undef CALL_
    }
    case CLSTM_KERNEL_DGRAD: {
#define CALL_(E) cell_dgrad<E>(p->ctx, cs, st)
      return DISPATCH_E(p->cfg.dtype, CALL_);
#undef CALL_
    }
    case CLSTM_KERNEL_WGRAD: {
#define CALL_(E) cell_wgrad<E>(p->ctx, cs, in, hslot(cs, step), 0, st)
      return DISPATCH_E(p->cfg.dtype, CALL_);
#undef CALL_
    }
    case 4: {  // experiment: weight-gradient on the side stream concurrently with a gate-gradient on `stream`
      Ctx& cx = p->ctx;
      if (!cx.side) return fail(CLSTM_ESTATE, "no side stream");
      CU_TRY(cudaEventRecord(cx.ev_fork, st));
      CU_TRY(cudaStreamWaitEvent(cx.side, cx.ev_fork, 0));
#define CALL_(E) cell_wgrad<E>(p->ctx, cs, in, hslot(cs, step), 0, cx.side, 1)
      RC_TRY(DISPATCH_E(p->cfg.dtype, CALL_));
#undef CALL_
#define CALL_(E) cell_gate_grad<E>(p->ctx, cs, gates, c_prev, c_next, cs.dh_own, p->dstack, nullptr, 0, st, 0)
      RC_TRY(DISPATCH_E(p->cfg.dtype, CALL_));
#undef CALL_
      CU_TRY(cudaEventRecord(cx.ev_w[0], cx.side));
      CU_TRY(cudaStreamWaitEvent(st, cx.ev_w[0], 0));
      return 0;
    }
    case CLSTM_KERNEL_DGRAD_FUSED: {  // dgrad of (cell, step) + gate gradient of the cell below at the same step
      if (cell < 1 || !cs.with_x || cs.g.CIP != 64 || !fuse_supported(p->ctx))
        return fail(CLSTM_EINVAL, "fused dgrad + gate-gradient is not available for this cell / shape");
      CellState& cn = p->cells[cell - 1];
      if (step >= cn.T) return fail(CLSTM_EINVAL, "step %d out of range for the consumer cell", step);
      // the kernel variant the default backward schedule launches (timing only: it runs on whatever the last backward left)
#define CALL_(E)                                                                                                \
  cell_dgrad_fused<E>(p->ctx, cs, cn, step, cn.dh_own, cs.dxb, nullptr, 0, st, 0x7fffffff, nullptr, nullptr, 0, \
                      state16_knobs_ok<E>(p->ctx) && p->ctx.knobs.hybrid_pct == 0 && p->L >= 2)
      return DISPATCH_E(p->cfg.dtype, CALL_);
#undef CALL_
    }
    default:
      return fail(CLSTM_EINVAL, "unknown kernel kind %d", kind);
  }
}

// ---------------------------------------------------------------------------- single cell
int clstm_cell_plan_create(int batch, int height, int width, int in_channels, int hidden, int kernel_h, int kernel_w,
                           int dtype, clstm_cell_plan_t** out) {
  if (!out) return fail(CLSTM_EINVAL, "null argument");
  *out = nullptr;
  RC_TRY(validate_common(batch, height, width, in_channels, hidden, kernel_h, kernel_w, dtype));
  clstm_cell_plan* p = new (std::nothrow) clstm_cell_plan();
  if (!p) return fail(CLSTM_EINVAL, "out of host memory");
  Ctx& ctx = p->ctx;
  DeviceInfo dev;
  if (get_device(&dev) == 0) ctx.dev = dev;
  ctx.geo = make_geo(batch, height, width);
  ctx.knobs.read();
  ctx.dtype = dtype;
  ctx.HP = pad_hidden(hidden);
  ctx.training = 1;
  ctx.grad_scale = 0.f;
  p->cin = in_channels, p->hid = hidden;
  init_cell(&p->cs, ctx, in_channels, hidden, kernel_h, kernel_w, 0, 1, 1, 0);  // x is plain user data
  p->cs.slots_h = 2, p->cs.slots_c = 2;
  carve_cell_plan(p, nullptr, nullptr);
  *out = p;
  return 0;
}

int clstm_cell_plan_destroy(clstm_cell_plan_t* plan) {
  delete plan;
  return 0;
}

size_t clstm_cell_plan_workspace_bytes(const clstm_cell_plan_t* plan) {
  return plan ? plan->saved_bytes + plan->scratch_bytes : 0;
}
size_t clstm_cell_plan_saved_bytes(const clstm_cell_plan_t* plan) { return plan ? plan->saved_bytes : 0; }
size_t clstm_cell_plan_scratch_bytes(const clstm_cell_plan_t* plan) { return plan ? plan->scratch_bytes : 0; }

int clstm_cell_plan_bind(clstm_cell_plan_t* p, void* workspace, size_t bytes, void* stream) {
  if (!p || !workspace) return fail(CLSTM_EINVAL, "null argument");
  if (bytes < p->saved_bytes + p->scratch_bytes)
    return fail(CLSTM_EINVAL, "workspace too small: %zu < %zu", bytes, p->saved_bytes + p->scratch_bytes);
  return clstm_cell_plan_bind_split(p, workspace, p->saved_bytes, static_cast<uint8_t*>(workspace) + p->saved_bytes,
                                    bytes - p->saved_bytes, stream);
}

int clstm_cell_plan_bind_split(clstm_cell_plan_t* p, void* saved, size_t saved_bytes, void* scratch,
                               size_t scratch_bytes, void* stream) {
  if (!p || !saved || !scratch) return fail(CLSTM_EINVAL, "null argument");
  if (saved_bytes < p->saved_bytes) return fail(CLSTM_EINVAL, "saved region too small: %zu < %zu", saved_bytes, p->saved_bytes);
  if (scratch_bytes < p->scratch_bytes)
    return fail(CLSTM_EINVAL, "scratch region too small: %zu < %zu", scratch_bytes, p->scratch_bytes);
  if (reinterpret_cast<uintptr_t>(saved) % 1024 || reinterpret_cast<uintptr_t>(scratch) % 1024)
    return fail(CLSTM_EINVAL, "workspace regions must be 1024-byte aligned");
  (void)stream;
  Ctx& ctx = p->ctx;
  if (ctx.dev.ordinal < 0 || !p->bound) {  // first bind: pin the device (later re-binds are host-only pointer updates)
    DeviceInfo dev;
    RC_TRY(get_device(&dev));
    if (ctx.dev.sms != 0 && ctx.dev.sms != dev.sms) return fail(CLSTM_ESTATE, "plan created for a different device");
    if (ctx.dev.sms == 0 && dev.sms != 148) return fail(CLSTM_ESTATE, "plan created without a device assumed 148 SMs");
    ctx.dev = dev;
  }
  carve_cell_plan(p, static_cast<uint8_t*>(saved), static_cast<uint8_t*>(scratch));
  const Geo& g = ctx.geo;
  RC_TRY(map_cell(p->cs, ctx));
  RC_TRY(make_map_act(&p->m_x128, ctx.dtype, p->xin, p->cs.g.CIP, g.W, g.H, g.B, g.BW, g.BH));
  RC_TRY(make_map_act(&p->m_x64, ctx.dtype, p->xin, p->cs.g.CIP, g.W, g.H, g.B, g.BW2, g.BH2));
  RC_TRY(make_map_act(&ctx.m_dz128, ctx.dtype, ctx.dz, 4 * ctx.HP, g.W, g.H, g.B, g.BW, g.BH));
  RC_TRY(make_map_act(&ctx.m_dz64, ctx.dtype, ctx.dz, 4 * ctx.HP, g.W, g.H, g.B, g.BW2, g.BH2));
  for (int i = 0; i < 2; ++i) ctx.m_dz128b[i] = ctx.m_dz128, ctx.m_dz64b[i] = ctx.m_dz64;
  p->bound = true;
  return 0;
}

int clstm_cell_forward(clstm_cell_plan_t* p, const float* x, const float* h_cur, const float* c_cur,
                       const float* weight, const float* bias, float* h_next, float* c_next, void* stream) {
  if (!p || !x || !h_cur || !c_cur || !weight) return fail(CLSTM_EINVAL, "null argument");
  if (!p->bound) return fail(CLSTM_ESTATE, "clstm_cell_forward before clstm_cell_plan_bind");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CALL_(E) cellplan_forward<E>(p, x, h_cur, c_cur, weight, bias, h_next, c_next, st)
  return DISPATCH_E(p->ctx.dtype, CALL_);
#undef CALL_
}

int clstm_cell_native_load(clstm_cell_plan_t* p, const float* x, const float* h, const float* c, const float* weight,
                           const float* bias, int reset_state, void* stream) {
  if (!p) return fail(CLSTM_EINVAL, "null argument");
  if (!p->bound) return fail(CLSTM_ESTATE, "clstm_cell_native_load before clstm_cell_plan_bind");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (reset_state) p->cur = 0;
#define CALL_(E) cellplan_native_load<E>(p, x, h, c, weight, bias, reset_state, st)
  RC_TRY(DISPATCH_E(p->ctx.dtype, CALL_));
#undef CALL_
  if (weight) p->native_ready = true;
  return 0;
}

int clstm_cell_native_step(clstm_cell_plan_t* p, void* stream) {
  if (!p) return fail(CLSTM_EINVAL, "null argument");
  if (!p->bound || !p->native_ready)
    return fail(CLSTM_ESTATE, "clstm_cell_native_step needs a bound plan and weights (clstm_cell_native_load)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CALL_(E) cellplan_native_step<E>(p, st)
  return DISPATCH_E(p->ctx.dtype, CALL_);
#undef CALL_
}

int clstm_cell_native_read(clstm_cell_plan_t* p, float* h_out, float* c_out, void* stream) {
  if (!p) return fail(CLSTM_EINVAL, "null argument");
  if (!p->bound) return fail(CLSTM_ESTATE, "clstm_cell_native_read before clstm_cell_plan_bind");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CALL_(E) cellplan_native_read<E>(p, h_out, c_out, st)
  return DISPATCH_E(p->ctx.dtype, CALL_);
#undef CALL_
}

int clstm_cell_backward(clstm_cell_plan_t* p, const float* dh_next, const float* dc_next, const float* weight,
                        float* dx, float* dh_cur, float* dc_cur, float* dweight, float* dbias, void* stream) {
  if (!p) return fail(CLSTM_EINVAL, "null argument");
  if (!p->bound) return fail(CLSTM_ESTATE, "clstm_cell_backward before clstm_cell_plan_bind");
  if (!p->forward_done) return fail(CLSTM_ESTATE, "clstm_cell_backward before any clstm_cell_forward on this plan");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CALL_(E) cellplan_backward<E>(p, dh_next, dc_next, weight, dx, dh_cur, dc_cur, dweight, dbias, st)
  return DISPATCH_E(p->ctx.dtype, CALL_);
#undef CALL_
}

// ---------------------------------------------------------------------------- fused loss
int clstm_mse_loss_grad(const float* y, const float* target, int batch, int channels, int t_out, int height, int width,
                        float* dy, float* partial, float* out, void* stream) {
  trace_begin(stream);
  if (!y || !target || !partial || !out) return fail(CLSTM_EINVAL, "null argument");
  if (batch < 1 || channels < 1 || t_out < 1 || t_out > 1024 || height < 1 || width < 1)
    return fail(CLSTM_EINVAL, "bad shape (t_out must be <= 1024)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long planes = static_cast<long long>(batch) * t_out * channels;
  if (planes > 0x7fffffffll) return fail(CLSTM_EINVAL, "too many planes");
  const int hw = height * width;
  const double n = static_cast<double>(planes) * hw;
  mse_loss_grad_kernel<<<static_cast<int>(planes), 256, 0, st>>>(y, target, dy, partial, batch, channels, t_out, hw,
                                                                  static_cast<float>(2.0 / n));
  RC_TRY(after_launch("mse_loss_grad_kernel"));
  mse_finalize_kernel<<<1, 1024, 0, st>>>(partial, out, batch, channels, t_out, static_cast<float>(1.0 / n),
                                          static_cast<float>(static_cast<double>(t_out) / n));
  return after_launch("mse_finalize_kernel");
}

}  // extern "C"
