// Whole-forward programs: the model-level entry points of the C ABI (include/vidsitu_b200.h, "clip program").
//
// A program is the recorded launch list of one forward of mdl_sf_base.SFBase.forward_encoder
// (vidsitu_code/mdl_sf_base.py:182-197: pack -> stems -> res stages + lateral convs -> pooled head -> proj_head)
// for a fixed batch size: conv plans, pools, the head's reductions, and the fork / join points of the two
// pathway streams (video_model_builder.py:124-131: the pathways only meet at the lateral convs).  It is built
// once (the planner appends ops in program order), then
//   vsb_program_run      replays it with ONE call: every launch on the caller's stream (lane 0) or on the
//                        program's side stream (lane 1), or - after vsb_program_capture - as one CUDA graph;
//   vsb_program_save     writes it to a file: every device pointer is stored as (memory region, offset), the
//                        contents of constant regions (packed weights, folded BatchNorm) travel with it;
//   vsb_program_load     rebuilds it in ANY host process (no Python): regions are placed in caller-provided (or
//                        library-allocated) device memory, constants uploaded, TMA descriptors re-encoded.
// Nothing here computes: every op calls the same entry point a per-op caller would.
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <string>
#include <vector>

#include "common.h"
#include "conv_plan.h"

using namespace vsb;

namespace {

enum OpKind : int {
  OP_CONV = 1,
  OP_BOTTLENECK = 2,
  OP_PACK = 3,
  OP_MAXPOOL = 4,
  OP_AVGPOOL = 5,
  OP_LINEAR = 6,
  OP_NL_ATTENTION = 7,
  OP_SCORE_ROWS = 8,
  OP_TRANSPOSE_PAD = 9,
  OP_SYNC = 10,
  OP_STEM_POOL = 11,
};

struct PackArgs {
  const uint8_t* frames;
  int n, t_in, h, w;
  int idx[64];
  int t_out;
  float mean[3], std[3];
  int reverse;
  void* out;
  int c_pad, out_w, x_off, dtype;
};
struct PoolArgs {
  const void* in;
  int n, t, h, w, c, in_pitch;
  void* out;
  int out_pitch, c_out, kt, kh, kw, st, sh, sw, pt, ph, pw, dtype;
};
struct AvgArgs {
  const void* in;
  int n, thw, c, in_pitch;
  float* feats;
  int feat_pitch, feat_off, dtype;
};
struct LinArgs {
  const float* x;
  int n, din;
  const float* w;
  const float* b;
  float* y;
  int dout, relu;
};
struct NlArgs {
  const void* theta;
  int theta_pitch;
  const void* phi;
  int phi_pitch;
  const void* g;
  int g_pitch;
  void* out;
  int out_pitch, n, tq, tk, c, softmax, dtype;
};
struct ScoreArgs {
  void* scores;
  long long rows;
  int valid, width, pitch, softmax;
};
struct TpArgs {
  const void* in;
  int in_pitch;
  void* out;
  int n, rows, cols, out_pitch;
};
struct SyncArgs {
  int from, to;
};

union OpArgs {
  vsb_conv_desc conv;
  vsb_bottleneck_desc bott;
  vsb_stem_pool_desc stem;
  PackArgs pack;
  PoolArgs pool;
  AvgArgs avg;
  LinArgs lin;
  NlArgs nl;
  ScoreArgs score;
  TpArgs tp;
  SyncArgs sync;
};

struct Op {
  int kind;
  int lane;
  char name[96];
  OpArgs a;
  // run-time state (never saved)
  const vsb_conv_plan* conv;
  const vsb_bottleneck_plan* bott;
  const vsb_stem_pool_plan* stem;
  bool owns_plan;
  cudaEvent_t ev;
};

struct Region {
  char name[64];
  int kind;
  char* ptr;
  unsigned long long bytes;
};

// byte offsets of the device-pointer fields of each op's argument block
#define PF(T, f) offsetof(T, f)
const size_t kConvPtrs[] = {PF(vsb_conv_desc, in),  PF(vsb_conv_desc, wgt),         PF(vsb_conv_desc, scale),
                            PF(vsb_conv_desc, bias), PF(vsb_conv_desc, residual),    PF(vsb_conv_desc, out),
                            PF(vsb_conv_desc, in2),  PF(vsb_conv_desc, tile_signal), PF(vsb_conv_desc, tile_wait)};
const size_t kBottPtrs[] = {PF(vsb_bottleneck_desc, x),  PF(vsb_bottleneck_desc, out), PF(vsb_bottleneck_desc, wa),
                            PF(vsb_bottleneck_desc, wb), PF(vsb_bottleneck_desc, wc),  PF(vsb_bottleneck_desc, sa),
                            PF(vsb_bottleneck_desc, ba), PF(vsb_bottleneck_desc, sb),  PF(vsb_bottleneck_desc, bb),
                            PF(vsb_bottleneck_desc, sc), PF(vsb_bottleneck_desc, bc)};
const size_t kStemPtrs[] = {PF(vsb_stem_pool_desc, in), PF(vsb_stem_pool_desc, wgt), PF(vsb_stem_pool_desc, scale),
                            PF(vsb_stem_pool_desc, bias), PF(vsb_stem_pool_desc, out)};
const size_t kPackPtrs[] = {PF(PackArgs, frames), PF(PackArgs, out)};
const size_t kPoolPtrs[] = {PF(PoolArgs, in), PF(PoolArgs, out)};
const size_t kAvgPtrs[] = {PF(AvgArgs, in), PF(AvgArgs, feats)};
const size_t kLinPtrs[] = {PF(LinArgs, x), PF(LinArgs, w), PF(LinArgs, b), PF(LinArgs, y)};
const size_t kNlPtrs[] = {PF(NlArgs, theta), PF(NlArgs, phi), PF(NlArgs, g), PF(NlArgs, out)};
const size_t kScorePtrs[] = {PF(ScoreArgs, scores)};
const size_t kTpPtrs[] = {PF(TpArgs, in), PF(TpArgs, out)};
#undef PF

struct PtrTable {
  const size_t* off;
  int n;
  size_t args_bytes;
};
#define TBL(arr, T) PtrTable{arr, (int)(sizeof(arr) / sizeof(arr[0])), sizeof(T)}
PtrTable ptr_table(int kind) {
  switch (kind) {
    case OP_CONV: return TBL(kConvPtrs, vsb_conv_desc);
    case OP_BOTTLENECK: return TBL(kBottPtrs, vsb_bottleneck_desc);
    case OP_STEM_POOL: return TBL(kStemPtrs, vsb_stem_pool_desc);
    case OP_PACK: return TBL(kPackPtrs, PackArgs);
    case OP_MAXPOOL: return TBL(kPoolPtrs, PoolArgs);
    case OP_AVGPOOL: return TBL(kAvgPtrs, AvgArgs);
    case OP_LINEAR: return TBL(kLinPtrs, LinArgs);
    case OP_NL_ATTENTION: return TBL(kNlPtrs, NlArgs);
    case OP_SCORE_ROWS: return TBL(kScorePtrs, ScoreArgs);
    case OP_TRANSPOSE_PAD: return TBL(kTpPtrs, TpArgs);
    case OP_SYNC: return PtrTable{nullptr, 0, sizeof(SyncArgs)};
    default: return PtrTable{nullptr, 0, 0};
  }
}
#undef TBL

constexpr char kMagic[8] = {'V', 'S', 'B', 'P', 'R', 'O', 'G', '1'};
constexpr unsigned long long kRegionAlign = 1024;  // TMA bases need 16 bytes; 1 KiB keeps every tensor sector-aligned

struct FileHeader {
  char magic[8];
  int abi;
  int n_regions;
  int n_ops;
  int reserved;
};
struct FileRegion {
  char name[64];
  int kind;
  int reserved;
  unsigned long long bytes;
};
struct FileOp {
  int kind;
  int lane;
  char name[96];
  unsigned long long args_bytes;
};

}  // namespace

struct vsb_program {
  std::vector<Op> ops;
  std::vector<Region> regions;
  cudaStream_t side = nullptr;
  cudaGraphExec_t graph = nullptr;
  char* owned_mem = nullptr;  // vsb_program_load with device_mem == NULL
  bool loaded = false;
};

namespace {

int run_op(const Op& op, cudaStream_t s) {
  void* st = (void*)s;
  switch (op.kind) {
    case OP_CONV: return vsb_conv3d_run(op.conv, st);
    case OP_BOTTLENECK: return vsb_bottleneck_run(op.bott, st);
    case OP_STEM_POOL: return vsb_stem_pool_run(op.stem, st);
    case OP_PACK: {
      const PackArgs& a = op.a.pack;
      return vsb_pack_frames(a.frames, a.n, a.t_in, a.h, a.w, a.idx, a.t_out, a.mean, a.std, a.reverse, a.out, a.c_pad,
                             a.out_w, a.x_off, a.dtype, st);
    }
    case OP_MAXPOOL: {
      const PoolArgs& a = op.a.pool;
      return vsb_maxpool3d(a.in, a.n, a.t, a.h, a.w, a.c, a.in_pitch, a.out, a.out_pitch, a.c_out, a.kt, a.kh, a.kw, a.st,
                           a.sh, a.sw, a.pt, a.ph, a.pw, a.dtype, st);
    }
    case OP_AVGPOOL: {
      const AvgArgs& a = op.a.avg;
      return vsb_global_avgpool(a.in, a.n, a.thw, a.c, a.in_pitch, a.feats, a.feat_pitch, a.feat_off, a.dtype, st);
    }
    case OP_LINEAR: {
      const LinArgs& a = op.a.lin;
      return vsb_linear(a.x, a.n, a.din, a.w, a.b, a.y, a.dout, a.relu, st);
    }
    case OP_NL_ATTENTION: {
      const NlArgs& a = op.a.nl;
      return vsb_nonlocal_attention(a.theta, a.theta_pitch, a.phi, a.phi_pitch, a.g, a.g_pitch, a.out, a.out_pitch, a.n,
                                    a.tq, a.tk, a.c, a.softmax, a.dtype, st);
    }
    case OP_SCORE_ROWS: {
      const ScoreArgs& a = op.a.score;
      return vsb_score_rows(a.scores, a.rows, a.valid, a.width, a.pitch, a.softmax, st);
    }
    case OP_TRANSPOSE_PAD: {
      const TpArgs& a = op.a.tp;
      return vsb_transpose_pad(a.in, a.in_pitch, a.out, a.n, a.rows, a.cols, a.out_pitch, st);
    }
    default: set_error("program: unknown op kind %d", op.kind); return VSB_ERR_INVALID;
  }
}

int run_ops(vsb_program* p, cudaStream_t origin) {
  cudaStream_t lanes[2] = {origin, nullptr};
  for (const Op& op : p->ops) {
    if (op.lane == 1 || (op.kind == OP_SYNC && (op.a.sync.from == 1 || op.a.sync.to == 1))) {
      if (!p->side) VSB_CHECK_CUDA(cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking));
      lanes[1] = p->side;
    }
    if (op.kind == OP_SYNC) {
      VSB_CHECK_CUDA(cudaEventRecord(op.ev, lanes[op.a.sync.from]));
      VSB_CHECK_CUDA(cudaStreamWaitEvent(lanes[op.a.sync.to], op.ev, 0));
      continue;
    }
    const int rc = run_op(op, lanes[op.lane]);
    if (rc != VSB_OK) return rc;  // the callee set the error text
  }
  return VSB_OK;
}

Op* push_op(vsb_program* p, int kind, int lane, const char* name) {
  if (!p || lane < 0 || lane > 1) {
    set_error("program: null program or lane %d outside {0, 1}", lane);
    return nullptr;
  }
  if (p->graph) {
    set_error("program: already captured; ops cannot be appended");
    return nullptr;
  }
  Op op;
  memset(&op, 0, sizeof(op));
  op.kind = kind;
  op.lane = lane;
  snprintf(op.name, sizeof(op.name), "%s", name ? name : "");
  p->ops.push_back(op);
  return &p->ops.back();
}

// (region index + 1) << 48 | offset for a device pointer; 0 for NULL
int relocate_out(const vsb_program* p, const void* ptr, const char* op_name, unsigned long long* out) {
  if (!ptr) {
    *out = 0;
    return VSB_OK;
  }
  const char* c = (const char*)ptr;
  for (size_t r = 0; r < p->regions.size(); ++r) {
    const Region& g = p->regions[r];
    if (c >= g.ptr && c < g.ptr + g.bytes) {
      *out = ((unsigned long long)(r + 1) << 48) | (unsigned long long)(c - g.ptr);
      return VSB_OK;
    }
  }
  set_error("program save: op '%s' points at %p, which no registered region contains", op_name, ptr);
  return VSB_ERR_INVALID;
}

bool read_exact(FILE* f, void* dst, size_t n) { return n == 0 || fread(dst, 1, n, f) == n; }
bool write_exact(FILE* f, const void* src, size_t n) { return n == 0 || fwrite(src, 1, n, f) == n; }

unsigned long long align_up(unsigned long long v) { return (v + kRegionAlign - 1) / kRegionAlign * kRegionAlign; }

}  // namespace

extern "C" int vsb_program_create(vsb_program** out) {
  VSB_CHECK_ARG(out, "null argument");
  *out = new (std::nothrow) vsb_program();
  VSB_CHECK_ARG(*out, "out of host memory");
  return VSB_OK;
}

extern "C" void vsb_program_destroy(vsb_program* p) {
  if (!p) return;
  if (p->graph) (void)cudaGraphExecDestroy(p->graph);
  for (Op& op : p->ops) {
    if (op.ev) (void)cudaEventDestroy(op.ev);
    if (op.owns_plan && op.conv) vsb_conv3d_plan_destroy(const_cast<vsb_conv_plan*>(op.conv));
    if (op.owns_plan && op.bott) vsb_bottleneck_plan_destroy(const_cast<vsb_bottleneck_plan*>(op.bott));
    if (op.owns_plan && op.stem) vsb_stem_pool_plan_destroy(const_cast<vsb_stem_pool_plan*>(op.stem));
  }
  if (p->side) (void)cudaStreamDestroy(p->side);
  if (p->owned_mem) (void)cudaFree(p->owned_mem);
  delete p;
}

extern "C" int vsb_program_add_region(vsb_program* p, const char* name, void* ptr, unsigned long long bytes, int kind) {
  VSB_CHECK_ARG(p && name && ptr && bytes > 0, "null program / name / pointer or empty region");
  VSB_CHECK_ARG(kind == VSB_REGION_CONST || kind == VSB_REGION_SCRATCH, "region kind must be VSB_REGION_CONST or VSB_REGION_SCRATCH");
  VSB_CHECK_ARG(strlen(name) < sizeof(Region::name), "region name longer than 63 characters");
  for (const Region& g : p->regions) {
    VSB_CHECK_ARG(strcmp(g.name, name) != 0, "region '%s' registered twice", name);
    VSB_CHECK_ARG((char*)ptr + bytes <= g.ptr || (char*)ptr >= g.ptr + g.bytes, "region '%s' overlaps region '%s'", name, g.name);
  }
  Region g;
  memset(&g, 0, sizeof(g));
  snprintf(g.name, sizeof(g.name), "%s", name);
  g.kind = kind;
  g.ptr = (char*)ptr;
  g.bytes = bytes;
  p->regions.push_back(g);
  return VSB_OK;
}

extern "C" int vsb_program_region(const vsb_program* p, const char* name, void** ptr, unsigned long long* bytes) {
  VSB_CHECK_ARG(p && name, "null argument");
  for (const Region& g : p->regions)
    if (strcmp(g.name, name) == 0) {
      if (ptr) *ptr = g.ptr;
      if (bytes) *bytes = g.bytes;
      return VSB_OK;
    }
  set_error("program has no region '%s'", name);
  return VSB_ERR_INVALID;
}

extern "C" int vsb_program_num_ops(const vsb_program* p) { return p ? (int)p->ops.size() : 0; }

extern "C" int vsb_program_num_launches(const vsb_program* p) {
  int n = 0;
  if (p)
    for (const Op& op : p->ops) n += op.kind == OP_SYNC ? 0 : (op.kind == OP_STEM_POOL ? 2 : 1);
  return n;
}

extern "C" unsigned long long vsb_program_device_bytes(const vsb_program* p) {
  unsigned long long total = 0;
  if (p)
    for (const Region& g : p->regions) total += align_up(g.bytes);
  return total;
}

extern "C" int vsb_program_add_conv(vsb_program* p, const vsb_conv_plan* plan, int lane, const char* name) {
  VSB_CHECK_ARG(plan, "null plan");
  Op* op = push_op(p, OP_CONV, lane, name);
  if (!op) return VSB_ERR_INVALID;
  op->a.conv = plan->desc;
  op->conv = plan;
  return VSB_OK;
}

extern "C" int vsb_program_add_bottleneck(vsb_program* p, const vsb_bottleneck_plan* plan, const vsb_bottleneck_desc* desc,
                                          int lane, const char* name) {
  VSB_CHECK_ARG(plan && desc, "null plan / desc");
  Op* op = push_op(p, OP_BOTTLENECK, lane, name);
  if (!op) return VSB_ERR_INVALID;
  op->a.bott = *desc;
  op->bott = plan;
  return VSB_OK;
}

extern "C" int vsb_program_add_stem_pool(vsb_program* p, const vsb_stem_pool_plan* plan, int lane, const char* name) {
  VSB_CHECK_ARG(plan, "null plan");
  vsb_stem_pool_desc d;
  int rc = vsb_stem_pool_plan_desc(plan, &d);
  if (rc != VSB_OK) return rc;
  Op* op = push_op(p, OP_STEM_POOL, lane, name);
  if (!op) return VSB_ERR_INVALID;
  op->a.stem = d;
  op->stem = plan;
  return VSB_OK;
}

extern "C" int vsb_program_add_pack_frames(vsb_program* p, const uint8_t* frames, int n, int t_in, int h, int w, const int* idx,
                                           int t_out, const float* mean3, const float* std3, int reverse_channels, void* out,
                                           int c_pad, int out_w, int x_off, int dtype, int lane, const char* name) {
  VSB_CHECK_ARG(frames && idx && mean3 && std3 && out && t_out > 0 && t_out <= 64, "null argument or t_out outside 1..64");
  Op* op = push_op(p, OP_PACK, lane, name);
  if (!op) return VSB_ERR_INVALID;
  PackArgs& a = op->a.pack;
  a.frames = frames;
  a.n = n, a.t_in = t_in, a.h = h, a.w = w, a.t_out = t_out;
  for (int i = 0; i < t_out; ++i) a.idx[i] = idx[i];
  for (int i = 0; i < 3; ++i) a.mean[i] = mean3[i], a.std[i] = std3[i];
  a.reverse = reverse_channels;
  a.out = out;
  a.c_pad = c_pad, a.out_w = out_w, a.x_off = x_off, a.dtype = dtype;
  return VSB_OK;
}

extern "C" int vsb_program_add_maxpool3d(vsb_program* p, const void* in, int n, int t, int h, int w, int c, int in_pitch,
                                         void* out, int out_pitch, int c_out, int kt, int kh, int kw, int st, int sh, int sw,
                                         int pt, int ph, int pw, int dtype, int lane, const char* name) {
  Op* op = push_op(p, OP_MAXPOOL, lane, name);
  if (!op) return VSB_ERR_INVALID;
  op->a.pool = PoolArgs{in, n, t, h, w, c, in_pitch, out, out_pitch, c_out, kt, kh, kw, st, sh, sw, pt, ph, pw, dtype};
  return VSB_OK;
}

extern "C" int vsb_program_add_global_avgpool(vsb_program* p, const void* in, int n, int thw, int c, int in_pitch, float* feats,
                                              int feat_pitch, int feat_off, int dtype, int lane, const char* name) {
  Op* op = push_op(p, OP_AVGPOOL, lane, name);
  if (!op) return VSB_ERR_INVALID;
  op->a.avg = AvgArgs{in, n, thw, c, in_pitch, feats, feat_pitch, feat_off, dtype};
  return VSB_OK;
}

extern "C" int vsb_program_add_linear(vsb_program* p, const float* x, int n, int din, const float* w, const float* b, float* y,
                                      int dout, int relu, int lane, const char* name) {
  Op* op = push_op(p, OP_LINEAR, lane, name);
  if (!op) return VSB_ERR_INVALID;
  op->a.lin = LinArgs{x, n, din, w, b, y, dout, relu};
  return VSB_OK;
}

extern "C" int vsb_program_add_nonlocal_attention(vsb_program* p, const void* theta, int theta_pitch, const void* phi,
                                                  int phi_pitch, const void* g, int g_pitch, void* out, int out_pitch, int n,
                                                  int tq, int tk, int c, int softmax, int dtype, int lane, const char* name) {
  Op* op = push_op(p, OP_NL_ATTENTION, lane, name);
  if (!op) return VSB_ERR_INVALID;
  op->a.nl = NlArgs{theta, theta_pitch, phi, phi_pitch, g, g_pitch, out, out_pitch, n, tq, tk, c, softmax, dtype};
  return VSB_OK;
}

extern "C" int vsb_program_add_score_rows(vsb_program* p, void* scores, long long rows, int valid, int width, int pitch,
                                          int softmax, int lane, const char* name) {
  Op* op = push_op(p, OP_SCORE_ROWS, lane, name);
  if (!op) return VSB_ERR_INVALID;
  op->a.score = ScoreArgs{scores, rows, valid, width, pitch, softmax};
  return VSB_OK;
}

extern "C" int vsb_program_add_transpose_pad(vsb_program* p, const void* in, int in_pitch, void* out, int n, int rows, int cols,
                                             int out_pitch, int lane, const char* name) {
  Op* op = push_op(p, OP_TRANSPOSE_PAD, lane, name);
  if (!op) return VSB_ERR_INVALID;
  op->a.tp = TpArgs{in, in_pitch, out, n, rows, cols, out_pitch};
  return VSB_OK;
}

extern "C" int vsb_program_add_sync(vsb_program* p, int from_lane, int to_lane) {
  VSB_CHECK_ARG(from_lane >= 0 && from_lane <= 1 && to_lane >= 0 && to_lane <= 1 && from_lane != to_lane,
                "sync joins lane 0 and lane 1");
  Op* op = push_op(p, OP_SYNC, 0, "sync");
  if (!op) return VSB_ERR_INVALID;
  op->a.sync = SyncArgs{from_lane, to_lane};
  VSB_CHECK_CUDA(cudaEventCreateWithFlags(&op->ev, cudaEventDisableTiming));
  return VSB_OK;
}

extern "C" int vsb_program_run(vsb_program* p, void* stream) {
  VSB_CHECK_ARG(p, "null program");
  cudaStream_t s = (cudaStream_t)stream;
  if (p->graph) {
    VSB_CHECK_CUDA(cudaGraphLaunch(p->graph, s));
    count_launch(vsb_program_num_launches(p));  // the graph's kernel nodes (run_ops counted them once, at capture)
    return VSB_OK;
  }
  return run_ops(p, s);
}

extern "C" int vsb_program_capture(vsb_program* p, void* stream) {
  VSB_CHECK_ARG(p, "null program");
  VSB_CHECK_ARG(!p->graph, "program already captured");
  cudaStream_t s = (cudaStream_t)stream;
  // one eager pass first: lazy function attributes / module loading must not happen inside the capture
  int rc = run_ops(p, s);
  if (rc != VSB_OK) return rc;
  VSB_CHECK_CUDA(cudaStreamSynchronize(s));
  if (p->side) VSB_CHECK_CUDA(cudaStreamSynchronize(p->side));
  // captured on a private stream (the caller's may be the legacy default stream, which cannot capture); the
  // instantiated graph launches on whatever stream vsb_program_run is given
  cudaStream_t cap = nullptr;
  VSB_CHECK_CUDA(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
  cudaError_t e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) {
    (void)cudaStreamDestroy(cap);
    set_error("cudaStreamBeginCapture failed: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return VSB_ERR_CUDA;
  }
  rc = run_ops(p, cap);
  cudaGraph_t g = nullptr;
  e = cudaStreamEndCapture(cap, &g);
  (void)cudaStreamDestroy(cap);
  if (rc != VSB_OK) {
    if (g) (void)cudaGraphDestroy(g);
    (void)cudaGetLastError();
    return rc;
  }
  if (e != cudaSuccess) {
    set_error("cudaStreamEndCapture failed: %s (does the program join lane 1 back into lane 0?)", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return VSB_ERR_CUDA;
  }
  e = cudaGraphInstantiate(&p->graph, g, 0);
  (void)cudaGraphDestroy(g);
  if (e != cudaSuccess) {
    p->graph = nullptr;
    set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return VSB_ERR_CUDA;
  }
  return VSB_OK;
}

extern "C" int vsb_program_save(const vsb_program* p, const char* path) {
  VSB_CHECK_ARG(p && path, "null argument");
  // relocate everything first: a pointer outside the registered regions must not leave half a file behind
  std::vector<std::vector<unsigned char>> blocks(p->ops.size());
  for (size_t i = 0; i < p->ops.size(); ++i) {
    const Op& op = p->ops[i];
    const PtrTable t = ptr_table(op.kind);
    VSB_CHECK_ARG(t.args_bytes > 0, "program save: unknown op kind %d", op.kind);
    blocks[i].assign((const unsigned char*)&op.a, (const unsigned char*)&op.a + t.args_bytes);
    for (int k = 0; k < t.n; ++k) {
      const void* ptr;
      memcpy(&ptr, blocks[i].data() + t.off[k], sizeof(ptr));
      unsigned long long rel;
      const int rc = relocate_out(p, ptr, op.name, &rel);
      if (rc != VSB_OK) return rc;
      memcpy(blocks[i].data() + t.off[k], &rel, sizeof(rel));
    }
  }
  FILE* f = fopen(path, "wb");
  VSB_CHECK_ARG(f, "cannot open '%s' for writing", path);
  FileHeader h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, kMagic, 8);
  h.abi = VSB_ABI_VERSION;
  h.n_regions = (int)p->regions.size();
  h.n_ops = (int)p->ops.size();
  bool ok = write_exact(f, &h, sizeof(h));
  std::vector<unsigned char> host;
  for (const Region& g : p->regions) {
    FileRegion fr;
    memset(&fr, 0, sizeof(fr));
    memcpy(fr.name, g.name, sizeof(fr.name));
    fr.kind = g.kind;
    fr.bytes = g.bytes;
    ok = ok && write_exact(f, &fr, sizeof(fr));
    if (g.kind == VSB_REGION_CONST) {
      host.resize(g.bytes);
      const cudaError_t e = cudaMemcpy(host.data(), g.ptr, g.bytes, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) {
        fclose(f);
        set_error("program save: reading region '%s' failed: %s", g.name, cudaGetErrorString(e));
        return VSB_ERR_CUDA;
      }
      ok = ok && write_exact(f, host.data(), g.bytes);
    }
  }
  for (size_t i = 0; i < p->ops.size(); ++i) {
    FileOp fo;
    memset(&fo, 0, sizeof(fo));
    fo.kind = p->ops[i].kind;
    fo.lane = p->ops[i].lane;
    memcpy(fo.name, p->ops[i].name, sizeof(fo.name));
    fo.args_bytes = blocks[i].size();
    ok = ok && write_exact(f, &fo, sizeof(fo)) && write_exact(f, blocks[i].data(), blocks[i].size());
  }
  ok = (fclose(f) == 0) && ok;
  VSB_CHECK_ARG(ok, "short write to '%s'", path);
  return VSB_OK;
}

extern "C" int vsb_program_file_device_bytes(const char* path, unsigned long long* bytes) {
  VSB_CHECK_ARG(path && bytes, "null argument");
  FILE* f = fopen(path, "rb");
  VSB_CHECK_ARG(f, "cannot open '%s'", path);
  FileHeader h;
  bool ok = read_exact(f, &h, sizeof(h)) && memcmp(h.magic, kMagic, 8) == 0;
  unsigned long long total = 0;
  for (int r = 0; ok && r < h.n_regions; ++r) {
    FileRegion fr;
    ok = read_exact(f, &fr, sizeof(fr));
    if (ok && fr.kind == VSB_REGION_CONST) ok = fseek(f, (long)fr.bytes, SEEK_CUR) == 0;
    if (ok) total += align_up(fr.bytes);
  }
  fclose(f);
  VSB_CHECK_ARG(ok, "'%s' is not a vidsitu_b200 program file", path);
  *bytes = total;
  return VSB_OK;
}

extern "C" int vsb_program_load(const char* path, void* device_mem, unsigned long long device_bytes, vsb_program** out) {
  VSB_CHECK_ARG(path && out, "null argument");
  *out = nullptr;
  unsigned long long need = 0;
  int rc = vsb_program_file_device_bytes(path, &need);
  if (rc != VSB_OK) return rc;
  VSB_CHECK_ARG(!device_mem || device_bytes >= need, "device memory too small: %llu bytes given, %llu needed", device_bytes, need);
  VSB_CHECK_ARG(!device_mem || ((uintptr_t)device_mem % kRegionAlign) == 0, "device memory must be 1 KiB aligned");
  FILE* f = fopen(path, "rb");
  VSB_CHECK_ARG(f, "cannot open '%s'", path);
  vsb_program* p = new (std::nothrow) vsb_program();
  if (!p) {
    fclose(f);
    set_error("out of host memory");
    return VSB_ERR_INVALID;
  }
  p->loaded = true;
#define LOAD_FAIL(code, ...)  \
  do {                        \
    set_error(__VA_ARGS__);   \
    fclose(f);                \
    vsb_program_destroy(p);   \
    return code;              \
  } while (0)
  FileHeader h;
  if (!read_exact(f, &h, sizeof(h)) || memcmp(h.magic, kMagic, 8) != 0) LOAD_FAIL(VSB_ERR_INVALID, "'%s' is not a vidsitu_b200 program file", path);
  if (h.abi != VSB_ABI_VERSION) LOAD_FAIL(VSB_ERR_INVALID, "program file was written by ABI %d, this library is ABI %d", h.abi, VSB_ABI_VERSION);
  char* base = (char*)device_mem;
  if (!base) {
    const cudaError_t e = cudaMalloc((void**)&p->owned_mem, need ? need : 1);
    if (e != cudaSuccess) LOAD_FAIL(VSB_ERR_CUDA, "cudaMalloc of %llu bytes failed: %s", need, cudaGetErrorString(e));
    base = p->owned_mem;
  }
  unsigned long long cursor = 0;
  std::vector<unsigned char> host;
  for (int r = 0; r < h.n_regions; ++r) {
    FileRegion fr;
    if (!read_exact(f, &fr, sizeof(fr))) LOAD_FAIL(VSB_ERR_INVALID, "truncated program file (region %d)", r);
    Region g;
    memset(&g, 0, sizeof(g));
    memcpy(g.name, fr.name, sizeof(g.name));
    g.name[sizeof(g.name) - 1] = 0;
    g.kind = fr.kind;
    g.bytes = fr.bytes;
    g.ptr = base + cursor;
    cursor += align_up(fr.bytes);
    cudaError_t e;
    if (fr.kind == VSB_REGION_CONST) {
      host.resize(fr.bytes);
      if (!read_exact(f, host.data(), fr.bytes)) LOAD_FAIL(VSB_ERR_INVALID, "truncated program file (contents of region '%s')", g.name);
      e = cudaMemcpy(g.ptr, host.data(), fr.bytes, cudaMemcpyHostToDevice);
    } else {
      // scratch starts as zeros: input rows keep a zero border that no kernel ever writes
      e = cudaMemset(g.ptr, 0, fr.bytes);
    }
    if (e != cudaSuccess) LOAD_FAIL(VSB_ERR_CUDA, "initialising region '%s' failed: %s", g.name, cudaGetErrorString(e));
    p->regions.push_back(g);
  }
  for (int i = 0; i < h.n_ops; ++i) {
    FileOp fo;
    if (!read_exact(f, &fo, sizeof(fo))) LOAD_FAIL(VSB_ERR_INVALID, "truncated program file (op %d)", i);
    const PtrTable t = ptr_table(fo.kind);
    if (t.args_bytes == 0 || fo.args_bytes != t.args_bytes) LOAD_FAIL(VSB_ERR_INVALID, "op %d: kind %d with %llu argument bytes is not known to this library", i, fo.kind, fo.args_bytes);
    Op op;
    memset(&op, 0, sizeof(op));
    op.kind = fo.kind;
    op.lane = fo.lane;
    memcpy(op.name, fo.name, sizeof(op.name));
    op.name[sizeof(op.name) - 1] = 0;
    if (!read_exact(f, &op.a, t.args_bytes)) LOAD_FAIL(VSB_ERR_INVALID, "truncated program file (arguments of op %d)", i);
    if (op.lane < 0 || op.lane > 1) LOAD_FAIL(VSB_ERR_INVALID, "op %d: lane %d", i, op.lane);
    for (int k = 0; k < t.n; ++k) {
      unsigned long long rel;
      memcpy(&rel, (unsigned char*)&op.a + t.off[k], sizeof(rel));
      void* ptr = nullptr;
      if (rel) {
        const unsigned long long r = (rel >> 48) - 1, off = rel & ((1ull << 48) - 1);
        if (r >= p->regions.size() || off >= p->regions[r].bytes) LOAD_FAIL(VSB_ERR_INVALID, "op '%s': pointer outside its region", op.name);
        ptr = p->regions[r].ptr + off;
      }
      memcpy((unsigned char*)&op.a + t.off[k], &ptr, sizeof(ptr));
    }
    if (op.kind == OP_SYNC) {
      if (op.a.sync.from < 0 || op.a.sync.from > 1 || op.a.sync.to < 0 || op.a.sync.to > 1) LOAD_FAIL(VSB_ERR_INVALID, "op %d: bad sync lanes", i);
      const cudaError_t e = cudaEventCreateWithFlags(&op.ev, cudaEventDisableTiming);
      if (e != cudaSuccess) LOAD_FAIL(VSB_ERR_CUDA, "cudaEventCreate failed: %s", cudaGetErrorString(e));
    }
    p->ops.push_back(op);
    Op& q = p->ops.back();
    if (q.kind == OP_CONV) {
      vsb_conv_plan* plan = nullptr;
      rc = vsb_conv3d_plan_create(&q.a.conv, &plan);
      if (rc != VSB_OK) {
        fclose(f);
        vsb_program_destroy(p);
        return rc;  // vsb_conv3d_plan_create set the error text
      }
      q.conv = plan;
      q.owns_plan = true;
    } else if (q.kind == OP_STEM_POOL) {
      vsb_stem_pool_plan* plan = nullptr;
      rc = vsb_stem_pool_plan_create(&q.a.stem, &plan);
      if (rc != VSB_OK) {
        fclose(f);
        vsb_program_destroy(p);
        return rc;
      }
      q.stem = plan;
      q.owns_plan = true;
    } else if (q.kind == OP_BOTTLENECK) {
      vsb_bottleneck_plan* plan = nullptr;
      rc = vsb_bottleneck_plan_create(&q.a.bott, &plan);
      if (rc != VSB_OK) {
        fclose(f);
        vsb_program_destroy(p);
        return rc;
      }
      q.bott = plan;
      q.owns_plan = true;
    }
  }
#undef LOAD_FAIL
  fclose(f);
  *out = p;
  return VSB_OK;
}
